"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol the headers declare, keeps the
reference's struct layout, and refuses (loudly) to compute without a GPU.  No compute calls are made here."""
import ctypes as C
import os
import re

import pytest

from _common import ROOT


@pytest.fixture(scope="module")
def lib():
    from quicked_b200 import build as qbuild
    qbuild.build()
    from quicked_b200 import capi
    return capi.load()


def declared_symbols():
    names = set()
    for hdr in ("quicked.h", "quicked_b200.h"):
        src = open(os.path.join(ROOT, "include", hdr)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        names |= set(re.findall(r"\b((?:quicked|qb200)_[a-z0-9_]+)\s*\(", src))
    return names


def test_exports_every_declared_symbol(lib):
    from quicked_b200 import capi
    decl = declared_symbols()
    assert {"quicked_new", "quicked_align", "quicked_free", "quicked_default_params", "quicked_check_error",
            "quicked_status_msg", "qb200_align_batch"} <= decl
    for name in decl:
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert set(capi.EXPORTS) == decl


def test_struct_layout_matches_reference():
    """reference quicked.h:43-67 on x86-64: params 48 B, aligner 72 B"""
    from quicked_b200 import capi
    assert C.sizeof(capi.Params) == 48 and C.sizeof(capi.Aligner) == 72
    assert capi.Params.hew_threshold.offset == 16 and capi.Params.only_score.offset == 32
    assert capi.Params.external_allocator.offset == 40
    assert capi.Aligner.cigar.offset == 16 and capi.Aligner.score.offset == 24 and capi.Aligner.timer.offset == 32


def test_defaults_and_messages(lib):
    from quicked_b200 import capi
    p = lib.quicked_default_params()                      # reference quicked.c:308-321
    assert (p.algo, p.bandwidth, p.window_size, p.overlap_size) == (0, 15, 9, 1)
    assert list(p.hew_threshold) == [40, 40] and list(p.hew_percentage) == [15, 15]
    assert not p.only_score and not p.force_scalar and not p.external_timer
    assert lib.quicked_status_msg(-4).decode() == "ERROR: Tried to align an empty sequence\n"
    assert lib.quicked_status_msg(1).decode() == "QuickEd finished without errors.\n"
    assert lib.quicked_check_error(-1) and not lib.quicked_check_error(1) and not lib.quicked_check_error(0)
    a = capi.Aligner()
    assert lib.quicked_new(C.byref(a), C.byref(p)) == 1   # QUICKED_WIP
    assert a.score == -1 and not a.cigar
    assert lib.quicked_align(C.byref(a), b"", 0, b"", 0) == -4          # before any device work
    p.algo = 17
    assert lib.quicked_align(C.byref(a), b"ACGT", 4, b"ACGT", 4) == -3  # QUICKED_UNKNOWN_ALGO
    assert lib.quicked_free(C.byref(a)) == 1


def test_no_cpu_fallback(lib):
    """without a GPU the product path must fail, not silently compute on the host"""
    if lib.qb200_device_count() > 0:
        pytest.skip("a GPU is visible")
    h = C.c_void_p()
    assert lib.qb200_create(C.byref(h), 0) == -100          # QB200_ERR_NO_DEVICE
    from quicked_b200 import capi
    p = lib.quicked_default_params()
    a = capi.Aligner()
    lib.quicked_new(C.byref(a), C.byref(p))
    assert lib.quicked_align(C.byref(a), b"ACGT", 4, b"ACTT", 4) == -1   # QUICKED_ERROR, message on stderr
    assert a.score == -1
    lib.quicked_free(C.byref(a))


def test_product_never_touches_oracle():
    """oracle/ is test infrastructure: nothing under quicked_b200/ or include/ may reference it"""
    for base in ("quicked_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                    txt = open(os.path.join(dp, f), errors="ignore").read()
                    assert "libqoracle" not in txt and "oracle.harness" not in txt and "quicked_oracle" not in txt, os.path.join(dp, f)


def test_native_generator_model():
    """seeded generate_dataset twin: text has exactly `length` bases, the pattern carries ceil(length*error) edits"""
    from quicked_b200 import generate_pairs_native
    seqs, po, pl, to, tl = generate_pairs_native(7, 200, 100, 0.05)
    assert (tl == 100).all() and (abs(pl - 100) <= 5).all()
    raw = seqs.tobytes()
    assert set(raw[po[0]:po[0] + pl[0]]) <= set(b"ACGT")
    seqs2, *_ = generate_pairs_native(7, 200, 100, 0.05)
    assert (seqs == seqs2).all()
    seqs3, *_ = generate_pairs_native(8, 200, 100, 0.05)
    assert (seqs != seqs3).any()


class _RefCigar(C.Structure):      # quicked_utils/include/cigar.h:33-47
    _fields_ = [("operations", C.c_char_p), ("cigar_buffer", C.POINTER(C.c_uint32)), ("cigar_length", C.c_int),
                ("max_operations", C.c_int), ("begin_offset", C.c_int), ("end_offset", C.c_int), ("score", C.c_int),
                ("end_v", C.c_int), ("end_h", C.c_int)]


def _expand(cigar):
    return "".join(op * int(n) for n, op in re.findall(r"(\d+)([MXID])", cigar))


SAM_GOLDEN = [("2M1X1M", "4M", "2=1X1="), ("3X2M1I", "1X4M1I", "3X2=1I"), ("1X", "1X", "1X"), ("5M", "5M", "5="),
              ("1M2X3D4I5M", "3M3D4I5M", "1=2X3D4I5="), ("2D1X1X3M", "2D5M", "2D2X3="), ("", "", "")]


def test_cigar_to_sam_golden():
    """qb200_cigar_to_sam on known answers (the first operation is never mapped: reference cigar.c:209)"""
    from quicked_b200.capi import cigar_to_sam
    for cig, plain, shown in SAM_GOLDEN:
        assert cigar_to_sam(cig, False) == plain
        assert cigar_to_sam(cig, True) == shown


def test_cigar_to_sam_matches_reference(reference):
    """against the unmodified reference's cigar_sprint_SAM_CIGAR (quicked_utils/src/cigar.c:504-529) on random CIGARs"""
    import numpy as np
    from quicked_b200.capi import cigar_to_sam
    ref = reference.lib
    ref.cigar_sprint_SAM_CIGAR.restype = C.c_int
    ref.cigar_sprint_SAM_CIGAR.argtypes = [C.c_char_p, C.c_int, C.POINTER(_RefCigar), C.c_bool]
    rng = np.random.default_rng(5)
    cases = [c for c, _, _ in SAM_GOLDEN if c]
    for _ in range(300):
        runs, last = [], ""
        for _ in range(int(rng.integers(1, 40))):
            op = str(rng.choice([o for o in "MMMXID" if o != last]))
            runs.append(f"{int(rng.integers(1, 30))}{op}")
            last = op
        cases.append("".join(runs))
    for cig in cases:
        ops = _expand(cig).encode()
        for show in (False, True):
            buf32 = (C.c_uint32 * (len(ops) + 1))()
            rc = _RefCigar(ops, buf32, 0, len(ops), 0, len(ops), 0, 0, 0)
            out = C.create_string_buffer(2 * len(ops) + 16)
            n = ref.cigar_sprint_SAM_CIGAR(out, len(out), C.byref(rc), show)
            assert cigar_to_sam(cig, show) == out.raw[:n].decode(), (cig, show)


def test_generator_twin_matches_the_checker_side_generator(lib):
    """qb200_generate_pairs_ex (product, host code) and oracle/datagen.c (checker side, used by bench.py --impl reference)
    are the same seeded model: byte-identical pairs, any slice of a job, with and without --indels."""
    import numpy as np
    import quicked_b200 as qb
    from oracle import harness
    for kw in (dict(first=0), dict(first=12345), dict(first=7, indels=(3, 40))):
        a = qb.generate_pairs_native(11, 64, 300, 0.1, **kw)
        b = harness.generate_pairs(11, 64, 300, 0.1, **kw)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
    whole = qb.generate_pairs_native(11, 100, 200, 0.2)
    part = qb.generate_pairs_native(11, 10, 200, 0.2, first=90)
    for i in range(10):
        j = 90 + i
        assert bytes(whole[0][whole[1][j]:whole[1][j] + whole[2][j]]) == bytes(part[0][part[1][i]:part[1][i] + part[2][i]])
        assert bytes(whole[0][whole[3][j]:whole[3][j] + whole[4][j]]) == bytes(part[0][part[3][i]:part[3][i] + part[4][i]])


def test_bench_helpers(oracle):
    """bench.py's parity gate: the CIGAR replay accepts the oracle's alignments and rejects corrupted ones; the
    mixed-batch deal covers every pair exactly once and re-packs sequences intact."""
    import sys
    import numpy as np
    sys.path.insert(0, ROOT)
    import bench
    from oracle import harness
    from quicked_b200.sharding import strided_deal
    s, po, pl, to, tl = harness.generate_pairs(5, 24, 500, 0.15)
    for i in range(24):
        p, t = bytes(s[po[i]:po[i] + pl[i]]), bytes(s[to[i]:to[i] + tl[i]])
        st, sc, cg = oracle.align(p, t)
        assert bench.replay_cigar(cg.encode(), p, t) == sc
        assert bench.replay_cigar(cg.encode().replace(b"M", b"X", 1), p, t) != sc
    seen = np.concatenate([strided_deal(pl, tl, r, 5) for r in range(5)])
    assert sorted(seen.tolist()) == list(range(24))
    sel = strided_deal(pl, tl, 2, 5)
    s2, po2, pl2, to2, tl2 = bench.gather_pairs(s, po, pl, to, tl, sel)
    for j, i in enumerate(sel):
        assert bytes(s2[po2[j]:po2[j] + pl2[j]]) == bytes(s[po[i]:po[i] + pl[i]])
        assert bytes(s2[to2[j]:to2[j] + tl2[j]]) == bytes(s[to[i]:to[i] + tl[i]])
    # roofline.traffic: per-kernel DRAM bytes of the committed ncu capture of the default command, keyed by bare kernel name
    import types
    traffic = bench.ncu_traffic(types.SimpleNamespace(workload="c3", algo="quicked"))
    assert traffic.get("k_band_tiles", 0) > 1e9 and traffic.get("k_windowed21_score", 0) > 1e9 and traffic.get("k_traceback_tiles", 0) > 1e9


def test_pack_2bit_round_trip(lib):
    """Host packer of the 2-bit upload format (qb200_pack_batch): unpacking the stream and patching the exception list
    (numpy restatement of the device kernels) gives back every sequence byte for byte — N, lower case, IUPAC and other
    bytes included; bytes between sequences are not kept."""
    import numpy as np
    from quicked_b200 import capi
    from quicked_b200.datagen import generate_pairs
    rng = np.random.default_rng(3)
    pairs = generate_pairs(50, 333, 0.1, seed=5) + generate_pairs(3, 5, 0.2, seed=6) + [("", "ACGT"), ("A", "")]
    pairs = [(p.encode() if isinstance(p, str) else p, t.encode() if isinstance(t, str) else t) for p, t in pairs]
    for i in range(0, len(pairs), 3):
        p, t = bytearray(pairs[i][0]), bytearray(pairs[i][1])
        for buf in (p, t):
            for k in rng.integers(0, max(1, len(buf)), size=4):
                if len(buf):
                    buf[k] = int(rng.choice(list(b"NnacgtRY*-\xff")))
        pairs[i] = (bytes(p), bytes(t))
    seqs, po, pl, to, tl = capi.pack_pairs(pairs)
    for threads in (1, 4):
        packed, ep, ec = capi.pack_2bit(seqs, po, pl, to, tl, threads=threads)
        assert np.all(np.diff(ep) > 0)
        back = capi.unpack_2bit(packed, int(seqs.size), ep, ec)
        for i, (p, t) in enumerate(pairs):
            assert bytes(back[po[i]:po[i] + pl[i]]) == p and bytes(back[to[i]:to[i] + tl[i]]) == t, i
        n_exc = sum(sum(c not in b"ACGT" for c in p) + sum(c not in b"ACGT" for c in t) for p, t in pairs)
        assert ep.size == n_exc
