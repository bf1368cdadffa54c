"""Drop-in evidence: the reference's UNMODIFIED callers — its ctest harness (tests/quicked_harness.c), its C examples, its
C++ binding (bindings/cpp/quicked.{hpp,cpp}) with the binding examples, and its pybind11 module
(bindings/python/quicked.cpp) with examples/bindings/basic.py — compiled from the reference tree against include/ and
linked to libquicked_b200.so (oracle/Makefile: refcallers -> oracle/_ref/callers/, git-ignored, shipped to the GPU box).

  * CPU: they build here (when /root/reference is present) and, without a GPU, fail loudly instead of answering.
  * GPU: they print exactly what the same sources print on the reference's own library (tests/golden/ref_callers.json,
    recorded by oracle/gen_callers_golden.py), including the two expectations the reference's ctest pins
    (tests/CMakeLists.txt:10-13): GATC/GATO -> score 1, ""/"" -> "ERROR: Tried to align an empty sequence".
"""
import json
import os
import subprocess
import sys

import pytest

from _common import GOLDEN, ROOT

CALLERS = os.path.join(ROOT, "oracle", "_ref", "callers")


def _build():
    from quicked_b200 import build as qbuild
    qbuild.build()
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "refcallers"], check=True)


def _run(name, argv):
    exe = name.split(":")[0]
    if exe == "binding_basic_py":
        return subprocess.run([sys.executable, os.path.join(CALLERS, "binding_basic.py")], capture_output=True, text=True,
                              env=dict(os.environ, PYTHONPATH=CALLERS), timeout=300)
    return subprocess.run([os.path.join(CALLERS, exe)] + argv, capture_output=True, text=True, timeout=300)


def test_reference_callers_build_and_refuse_without_gpu():
    if not os.path.isdir("/root/reference/examples") and not os.path.isdir(CALLERS):
        pytest.skip("reference tree absent and no prebuilt callers")
    _build()
    gold = json.load(open(os.path.join(GOLDEN, "ref_callers.json")))
    for name in gold:
        exe = name.split(":")[0]
        path = os.path.join(CALLERS, "binding_basic.py" if exe == "binding_basic_py" else exe)
        assert os.path.exists(path), f"{path} was not built"
    assert any(f.startswith("pyquicked") and f.endswith(".so") for f in os.listdir(CALLERS))
    from quicked_b200 import load
    if load().qb200_device_count() > 0:
        return
    # no GPU: the library must say so and the caller must see an error status, never a made-up answer
    r = _run("quicked_harness:nonDNA", ["GATC", "GATO", "1"])
    assert r.returncode != 0 and "no CUDA device" in r.stderr and "Got score" not in r.stdout
    r = _run("quicked_harness:empty", ["", ""])          # the empty-sequence check comes before any device work
    assert r.returncode != 0 and "ERROR: Tried to align an empty sequence" in r.stderr


@pytest.mark.gpu
def test_reference_callers_give_the_reference_output():
    if not os.path.isdir(CALLERS):
        pytest.skip("oracle/_ref/callers not built (reference tree absent when the snapshot was made)")
    gold = json.load(open(os.path.join(GOLDEN, "ref_callers.json")))
    for name, exp in sorted(gold.items()):
        r = _run(name, exp["argv"])
        assert (r.returncode, r.stdout) == (exp["rc"], exp["stdout"]), (name, r.returncode, r.stdout, r.stderr)
        assert r.stderr == exp["stderr"], (name, r.stderr)
