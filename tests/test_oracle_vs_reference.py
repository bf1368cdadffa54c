"""CPU tests (no GPU): pin the oracle restatement (oracle/quicked_oracle.c) to

  (a) the reference's own known answers (tests/CMakeLists.txt:10-13, examples),
  (b) committed golden vectors dumped from the unmodified reference (tests/golden/, oracle/gen_golden.py),
  (c) the unmodified reference itself (oracle/_ref/libquicked_ref.so) on fresh seeded inputs, when it is present.
"""
import os

import numpy as np
import pytest

from quicked_b200.datagen import generate_pairs, read_seq_file
from _common import GOLDEN, expand_rle, golden_kw, load_golden, replay, sha


# ---------------------------------------------------------------- (a) known answers
def test_known_answers(oracle):
    assert oracle.align("GATC", "GATO")[:2] == (1, 1)                 # tests/CMakeLists.txt:13 (non-ACGT -> code 4)
    for algo in (0, 1, 2):
        assert oracle.align("ACGT", "ACTT", algo=algo) == (1, 1, "2M1X1M")   # examples/*.c, QUICKED_WIP
    assert oracle.align("ACGT", "ACTT", algo=3) == (0, 1, "2M1X1M")   # HIRSCHBERG returns QUICKED_OK
    st, sc, cg = oracle.align("", "")
    assert st == -4 and cg is None                                     # tests/CMakeLists.txt:10-11
    assert oracle.status_msg(st).startswith("ERROR: Tried to align an empty sequence")
    assert oracle.align("ACGT", "")[0] == -4 and oracle.align("", "ACGT")[0] == -4
    assert oracle.align("ACGT", "ACGT", algo=7)[0] == -3               # QUICKED_UNKNOWN_ALGO


# ---------------------------------------------------------------- (b) golden vectors
def test_golden_explicit(oracle):
    gold = load_golden("golden_explicit.json")
    n = 0
    for case in gold["cases"]:
        for aname, exp in case["out"].items():
            got = oracle.align(case["pattern"], case["text"], **golden_kw(gold, aname))
            assert got == (exp["status"], exp["score"], exp["cigar"]), (aname, case["pattern"], case["text"])
            n += 1
    assert n > 300


@pytest.mark.parametrize("set_idx", range(8))
def test_golden_seeded(oracle, set_idx):
    gold = load_golden("golden_seeded.json")
    s = gold["sets"][set_idx]
    pairs = generate_pairs(s["num"], s["length"], s["error"], seed=s["seed"],
                           indels=tuple(s["indels"]) if s["indels"] else None)
    assert sha("".join(p.decode() + "|" + t.decode() + "\n" for p, t in pairs)) == s["inputs_sha1"], \
        "seeded generator drifted: regenerate goldens"
    for aname, rows in s["out"].items():
        kw = golden_kw(gold, aname)
        for (p, t), exp in zip(pairs, rows):
            if exp is None:       # the reference is undefined on this input (uninitialised read)
                continue
            st, sc, cg = oracle.align(p, t, **kw)
            assert [st, sc, sha(cg)] == exp, (s["name"], aname)


def test_golden_ont_pair(oracle):
    """tests/CMakeLists.txt:32 (test_MiniION_align_benchmark): QUICKED score == edlib on the real ONT pair.
    The unmodified reference built here gives 39743 on the 508596 x 505792 pair (SURVEY quotes 39740)."""
    gold = load_golden("golden_seeded.json")["ont"]
    p, t = read_seq_file(os.path.join(GOLDEN, "ONT.MiniION.1.seq"))[0]
    assert (len(p), len(t)) == (gold["m"], gold["n"])
    st, sc, cg = oracle.align(p, t)
    assert (st, sc, sha(cg)) == (gold["status"], gold["score"], gold["cigar_sha1"])
    assert replay(expand_rle(cg), p.decode(), t.decode()) == sc


# ---------------------------------------------------------------- (c) live against the reference
FAMILIES = [(60, 0.05, 30), (100, 0.05, 120), (100, 0.30, 60), (300, 0.15, 60), (1000, 0.10, 60), (1000, 0.30, 30),
            (3000, 0.20, 12), (10000, 0.20, 4), (30000, 0.15, 1)]


@pytest.mark.parametrize("length,error,num", FAMILIES)
def test_all_algos_match_reference(oracle, reference, length, error, num):
    for p, t in generate_pairs(num, length, error, seed=1000 + length):
        for algo in (0, 1, 2, 3):
            for fs in (False, True):
                assert oracle.align(p, t, algo=algo, force_scalar=fs) == reference.align(p, t, algo=algo, force_scalar=fs)


def test_bound_stages_2_and_3_match_reference(oracle, reference):
    stages = {1: 0, 2: 0, 3: 0}
    for length, error, num, indels in [(3000, 0.05, 20, (4, 200)), (10000, 0.1, 10, (4, 400)), (20000, 0.02, 4, (6, 600))]:
        for p, t in generate_pairs(num, length, error, seed=77, indels=indels):
            got = oracle.align(p, t, full=True)
            stages[got[3]["stage"]] += 1
            if got[3]["ref_undefined"]:
                continue
            assert got[:3] == reference.align(p, t)
    assert stages[2] + stages[3] > 10 and stages[3] > 5      # the test really exercises WindowEd(L) and band doubling


def test_params_sweep_matches_reference(oracle, reference):
    for length, error, num in [(1000, 0.2, 8), (5000, 0.2, 3), (200, 0.1, 10)]:
        for p, t in generate_pairs(num, length, error, seed=11):
            for bw in (1, 3, 5, 10, 20, 40):
                for only_score in (False, True):
                    kw = dict(algo=2, bandwidth=bw, only_score=only_score)
                    if oracle.align(p, t, full=True, **kw)[3]["ref_undefined"]:
                        continue
                    assert oracle.align(p, t, **kw) == reference.align(p, t, **kw), kw
                assert oracle.align(p, t, algo=3, bandwidth=bw) == reference.align(p, t, algo=3, bandwidth=bw)
            for W, O in [(2, 1), (3, 1), (9, 1), (4, 2), (9, 3), (1, 0), (16, 1)]:
                for only_score in (False, True):
                    for fs in (False, True):
                        kw = dict(algo=1, window_size=W, overlap_size=O, only_score=only_score, force_scalar=fs)
                        assert oracle.align(p, t, **kw) == reference.align(p, t, **kw), kw


def test_non_acgt_and_case_match_reference(oracle, reference):
    rng = np.random.default_rng(5)
    for _ in range(150):
        m = int(rng.integers(200, 700))
        n = m + int(rng.integers(-20, 21))        # the reference corrupts its heap on very ragged pairs (e.g. 39 x 210)
        p = bytes(rng.choice(list(b"ACGTN"), size=m).astype(np.uint8))
        t = bytes(rng.choice(list(b"ACGTNacgt"), size=n).astype(np.uint8))
        for algo in (0, 1, 2, 3):
            assert oracle.align(p, t, algo=algo) == reference.align(p, t, algo=algo)


def test_hirschberg_split_matches_reference(oracle, reference):
    p, t = generate_pairs(1, 100000, 0.2, seed=3)[0]
    got = oracle.align(p, t, full=True)
    assert got[3]["splits"] >= 7
    assert got[:3] == reference.align(p, t)
    assert oracle.align(p, t, algo=3, bandwidth=20) == reference.align(p, t, algo=3, bandwidth=20)


def test_oracle_cigars_replay(oracle):
    for length, error in [(100, 0.05), (1000, 0.1), (4000, 0.25)]:
        for p, t in generate_pairs(5, length, error, seed=9):
            for algo in (0, 1, 2, 3):
                st, sc, ops = oracle.align_ops(p, t, algo=algo)
                assert replay(ops, p.decode(), t.decode()) == sc


def test_native_batch_drivers_agree(oracle, reference):
    """the native batch drivers used for the CPU baseline (reference: oracle/ref_batch.c, port: qo_batch_align) return
    the same scores as pair-by-pair calls"""
    import numpy as np
    from oracle import harness
    from quicked_b200 import pack_pairs
    pairs = generate_pairs(64, 800, 0.12, seed=5) + generate_pairs(64, 100, 0.05, seed=6)
    seqs, po, pl, to, tl = pack_pairs(pairs)
    kind, nbytes, scores = harness.cpu_batch_align(seqs, po, pl, to, tl, 4, algo=0, want_scores=True)
    assert kind == "reference" and nbytes > 0
    exp = [oracle.align(p, t)[1] for p, t in pairs]
    assert scores.tolist() == exp
    # the port's driver
    import ctypes as C
    o = harness.Oracle()
    o.lib.qo_batch_align.restype = C.c_int64
    sc2 = np.zeros(len(pairs), np.int32)
    p = o.params(algo=0)
    o.lib.qo_batch_align(C.c_void_p(seqs.ctypes.data), C.c_void_p(po.ctypes.data), C.c_void_p(pl.ctypes.data), C.c_void_p(to.ctypes.data),
                         C.c_void_p(tl.ctypes.data), C.c_int64(len(pairs)), C.c_int(3), C.byref(p), C.c_void_p(sc2.ctypes.data))
    assert sc2.tolist() == exp
