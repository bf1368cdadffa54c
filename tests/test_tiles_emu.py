"""CPU test of the tile-dataflow BandEd kernels (quicked_b200/csrc/qb_tiles.cuh, qb_tiletrace.cuh).

The scheduler (band decisions, round planning), the tile update and the tile-record traceback are __host__ __device__
code; tests/emu/tile_emu.cu drives exactly that code on the CPU — task slots, packing of tiles onto lanes, the barriers of
a pass replaced by loops in shuffled order — and compares with the oracle: score-only passes (score, lower/higher block,
exported Pv/Mv state, scores[] and the word-step count) and full-matrix leaves (live ranges, every tile record against
the oracle's stored matrix column by column, and the walk's op string).  Too-narrow bands, ragged pairs, partial and
reversed passes, odd characters included."""
import os
import shutil
import subprocess

import pytest

from _common import ROOT


@pytest.mark.skipif(shutil.which("nvcc") is None and not os.path.exists("/usr/local/cuda/bin/nvcc"), reason="nvcc not available")
def test_tile_emulation_matches_oracle(oracle):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    exe = os.path.join(ROOT, "tests", "emu", "tile_emu")
    src = os.path.join(ROOT, "tests", "emu", "tile_emu.cu")
    odir = os.path.join(ROOT, "oracle")
    deps = [src] + [os.path.join(ROOT, "quicked_b200", "csrc", f) for f in ("qb_tiles.cuh", "qb_tiletrace.cuh", "qb_common.cuh", "qb_traceback.cuh")]
    if not os.path.exists(exe) or any(os.path.getmtime(d) > os.path.getmtime(exe) for d in deps):
        r = subprocess.run([nvcc, "-O2", "-std=c++17", "-w", "-x", "cu", "-o", exe, src, f"-L{odir}", "-lqoracle", "-Xlinker", f"-rpath,{odir}"],
                           capture_output=True, text=True)
        assert r.returncode == 0, r.stderr[-2000:]
    r = subprocess.run([exe, "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "tile emulation OK" in r.stdout, (r.stdout + r.stderr)[-3000:]
    assert "0 punts" in r.stderr            # the fill never gave a task up on these inputs
