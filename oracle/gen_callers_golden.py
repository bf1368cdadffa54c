"""oracle/gen_callers_golden.py — TEST INFRASTRUCTURE.  Runs the reference's own callers (its ctest harness, its C
examples, its C++ binding examples, its pybind11 example) built against the UNMODIFIED reference library
(oracle/_ref/libquicked_ref.so) and records what they print in tests/golden/ref_callers.json.
tests/test_reference_callers.py then runs the SAME unmodified sources built against include/ + libquicked_b200.so
(oracle/Makefile: refcallers) on the GPU box and expects identical output.  Needs /root/reference; run from the repo root:
    python oracle/gen_callers_golden.py
"""
import json
import os
import subprocess
import sys
import sysconfig
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("REF", "/root/reference")
REFLIB = os.path.join(ROOT, "oracle", "_ref")

CASES = {          # name -> argv tail (the examples take none)
    "quicked_harness:nonDNA": ["GATC", "GATO", "1"],          # tests/CMakeLists.txt:13
    "quicked_harness:empty": ["", ""],                        # tests/CMakeLists.txt:10-11
    "quicked_harness:acgt": ["ACGT", "ACTT", "1"],
    "quicked_harness:wrong": ["ACGTACGTAC", "ACGTTCGTAC", "3"],
    "example_basic": [], "example_banded": [], "example_banded_score": [], "example_windowed": [],
    "example_windowed_score": [], "example_hirschberg": [], "binding_basic_cpp": [], "binding_params_cpp": [],
    "binding_basic_py": [],
}


def main():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle"), "ref"], check=True)
    inc = [f"-I{REF}", f"-I{REF}/quicked", f"-I{REF}/quicked/include", f"-I{REF}/quicked_utils/include"]
    link = [f"-L{REFLIB}", "-lquicked_ref", f"-Wl,-rpath,{REFLIB}", "-lm"]
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        def cc(cmd):
            subprocess.run(cmd, check=True)
        cc(["gcc", "-O2", "-w"] + inc + ["-o", f"{tmp}/quicked_harness", f"{REF}/tests/quicked_harness.c"] + link)
        for ex in ("basic", "banded", "banded_score", "windowed", "windowed_score", "hirschberg"):
            cc(["gcc", "-O2", "-w"] + inc + ["-o", f"{tmp}/example_{ex}", f"{REF}/examples/{ex}.c"] + link)
        for ex in ("basic", "params"):
            cc(["g++", "-O2", "-w", "-std=c++11", f"-I{REF}/bindings/cpp"] + inc + ["-o", f"{tmp}/binding_{ex}_cpp",
                f"{REF}/examples/bindings/{ex}.cpp", f"{REF}/bindings/cpp/quicked.cpp"] + link)
        ext = sysconfig.get_config_var("EXT_SUFFIX")
        cc(["g++", "-O2", "-w", "-std=c++11", "-shared", "-fPIC", f"-I{REF}/bindings/cpp", f"-I{REF}/bindings/python/pybind11/include",
            f"-I{sysconfig.get_paths()['include']}"] + inc + ["-o", f"{tmp}/pyquicked{ext}", f"{REF}/bindings/python/quicked.cpp",
            f"{REF}/bindings/cpp/quicked.cpp"] + link)
        for name, argv in CASES.items():
            exe = name.split(":")[0]
            if exe == "binding_basic_py":
                cmd = [sys.executable, f"{REF}/examples/bindings/basic.py"]
                env = dict(os.environ, PYTHONPATH=tmp)
            else:
                cmd, env = [f"{tmp}/{exe}"] + argv, None
            r = subprocess.run(cmd, capture_output=True, text=True, env=env)
            out[name] = {"argv": argv, "rc": r.returncode, "stdout": r.stdout, "stderr": r.stderr}
            print(name, r.returncode, repr(r.stdout[-80:]), repr(r.stderr[-80:]))
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "ref_callers.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
