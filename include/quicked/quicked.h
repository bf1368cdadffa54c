/* include/quicked/quicked.h — the path the reference's C++ binding uses (bindings/cpp/quicked.hpp:33 includes
 * "quicked/quicked.h"): lets bindings/cpp/quicked.{hpp,cpp} and the pybind11 module build unchanged with -I include. */
#include "../quicked.h"
