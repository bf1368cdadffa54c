"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C-ABI
(libquicked_b200.so via quicked_b200.capi) and is compared bit-exactly with the CPU oracle on the same
seeded inputs, and with the committed golden vectors dumped from the unmodified reference."""
import os

import numpy as np
import pytest

from quicked_b200.datagen import generate_pairs
from _common import GOLDEN, expand_rle, golden_kw, load_golden, replay, sha

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from quicked_b200 import BatchAligner, load
    lib = load()
    assert lib.qb200_device_count() > 0, "no CUDA device visible: the GPU tests cannot run"
    a = BatchAligner(device=0)
    yield a
    a.close()


def check_against_oracle(gpu, oracle, pairs, allow_unimplemented=False, **kw):
    got = gpu.align(pairs, **kw)
    n_checked = 0
    for (p, t), g in zip(pairs, got):
        if allow_unimplemented and g[0] == -10:
            continue
        exp = oracle.align(p, t, **kw)
        assert g == exp, (kw, len(p), len(t), g[:2], exp[:2])
        n_checked += 1
    return n_checked


@pytest.mark.parametrize("length,error,num", [(100, 0.05, 400), (1000, 0.10, 200), (300, 0.25, 100), (3000, 0.2, 40),
                                              (10000, 0.2, 16), (64, 0.1, 64), (130, 0.02, 64)])
def test_quicked_matches_oracle(gpu, oracle, length, error, num):
    pairs = generate_pairs(num, length, error, seed=2000 + length)
    assert check_against_oracle(gpu, oracle, pairs, algo=0) == num
    assert check_against_oracle(gpu, oracle, pairs, algo=0, force_scalar=True) == num


@pytest.mark.parametrize("bandwidth", [5, 15, 20, 40, 150, 400])
def test_banded_matches_oracle(gpu, oracle, bandwidth):
    for length, error, num in [(100, 0.05, 100), (1000, 0.1, 60), (10000, 0.2, 8), (40, 0.1, 60), (150, 0.3, 60)]:
        if bandwidth > 100 and length > 1000:
            continue
        pairs = generate_pairs(num, length, error, seed=3000 + length)
        check_against_oracle(gpu, oracle, pairs, algo=2, bandwidth=bandwidth)


def test_hirschberg_no_split_matches_oracle(gpu, oracle):
    for length, error, num in [(100, 0.05, 100), (1000, 0.1, 60), (10000, 0.2, 8)]:
        pairs = generate_pairs(num, length, error, seed=4000 + length)
        check_against_oracle(gpu, oracle, pairs, algo=3, bandwidth=20)


def test_known_answers(gpu):
    got = gpu.align([("GATC", "GATO"), ("ACGT", "ACTT"), ("", ""), ("ACGT", ""), ("A", "A")])
    assert got[0][:2] == (1, 1)
    assert got[1] == (1, 1, "2M1X1M")
    assert got[2][0] == -4 and got[3][0] == -4
    assert got[4] == (1, 0, "1M")
    assert all(g[0] == -3 for g in gpu.align([("ACGT", "ACGT")], algo=9))


def test_golden_explicit(gpu):
    gold = load_golden("golden_explicit.json")
    pairs = [(c["pattern"], c["text"]) for c in gold["cases"]]
    for aname in gold["params"]:
        got = gpu.align(pairs, **golden_kw(gold, aname))
        for c, g in zip(gold["cases"], got):
            if aname in c["out"]:
                e = c["out"][aname]
                assert g == (e["status"], e["score"], e["cigar"]), (aname, c["pattern"], c["text"])


@pytest.mark.parametrize("set_idx", range(8))
def test_golden_seeded(gpu, set_idx):
    """every committed golden set (dumped from the unmodified reference), every algorithm / parameter variant"""
    gold = load_golden("golden_seeded.json")
    s = gold["sets"][set_idx]
    pairs = generate_pairs(s["num"], s["length"], s["error"], seed=s["seed"], indels=tuple(s["indels"]) if s["indels"] else None)
    for aname in gold["params"]:
        got = gpu.align(pairs, **golden_kw(gold, aname))
        for g, exp in zip(got, s["out"][aname]):
            if exp is None:          # the reference itself is undefined on this input
                continue
            assert [g[0], g[1], sha(g[2])] == exp, (s["name"], aname)


def test_golden_ont_pair(gpu):
    """reference tests/CMakeLists.txt:32 — the real 508 kbp MinION pair: stages 2-3, 25 Hirschberg splits"""
    gold = load_golden("golden_seeded.json")["ont"]
    from quicked_b200.datagen import read_seq_file
    p, t = read_seq_file(os.path.join(GOLDEN, "ONT.MiniION.1.seq"))[0]
    st, sc, cg = gpu.align([(p, t)])[0]
    assert (st, sc, sha(cg)) == (gold["status"], gold["score"], gold["cigar_sha1"])
    stats = gpu.stats()
    assert stats["hirschberg_splits"] >= 20 and stats["pairs_stage3"] == 1


def test_bound_stages_2_and_3(gpu, oracle):
    n2 = n3 = 0
    for length, error, num, indels in [(3000, 0.05, 24, (4, 200)), (10000, 0.1, 10, (4, 400)), (1000, 0.2, 24, (2, 150))]:
        pairs = generate_pairs(num, length, error, seed=77, indels=indels)
        for fs in (False, True):
            got = gpu.align(pairs, algo=0, force_scalar=fs)
            st = gpu.stats()
            n2 += st["pairs_stage2"]; n3 += st["pairs_stage3"]
            for (p, t), g in zip(pairs, got):
                assert g == oracle.align(p, t, algo=0, force_scalar=fs), (length, fs)
    assert n2 > 20 and n3 > 10


def test_hirschberg_splits(gpu, oracle):
    pairs = generate_pairs(2, 100000, 0.2, seed=3) + generate_pairs(3, 30000, 0.15, seed=4)
    for kw in (dict(algo=0), dict(algo=3, bandwidth=20), dict(algo=3, bandwidth=40)):
        got = gpu.align(pairs, **kw)
        for (p, t), g in zip(pairs, got):
            assert g == oracle.align(p, t, **kw), kw
    assert gpu.stats()["hirschberg_splits"] > 0


@pytest.mark.parametrize("W,O", [(2, 1), (3, 1), (9, 1), (4, 2), (9, 3), (1, 0), (16, 1), (2, 0)])
def test_windowed_matches_oracle(gpu, oracle, W, O):
    for length, error, num in [(200, 0.1, 30), (1000, 0.2, 20), (5000, 0.2, 4)]:
        pairs = generate_pairs(num, length, error, seed=11)
        for only_score in (False, True):
            for fs in (False, True):
                kw = dict(algo=1, window_size=W, overlap_size=O, only_score=only_score, force_scalar=fs)
                got = gpu.align(pairs, **kw)
                for (p, t), g in zip(pairs, got):
                    assert g == oracle.align(p, t, **kw), kw


def test_banded_only_score(gpu, oracle):
    for length, error, num in [(200, 0.1, 30), (1000, 0.2, 20), (10000, 0.2, 6)]:
        pairs = generate_pairs(num, length, error, seed=13)
        for bw in (1, 5, 10, 20, 40):
            kw = dict(algo=2, bandwidth=bw, only_score=True)
            got = gpu.align(pairs, **kw)
            for (p, t), g in zip(pairs, got):
                assert g == oracle.align(p, t, **kw), kw


def test_only_score_quicked_is_true_distance(gpu, oracle):
    """only_score for QUICKED/HIRSCHBERG is uninitialised in the reference; we return the
    distance of the traced alignment and no CIGAR."""
    pairs = generate_pairs(50, 500, 0.1, seed=17)
    got = gpu.align(pairs, algo=0, only_score=True)
    for (p, t), g in zip(pairs, got):
        exp = oracle.align(p, t, algo=0)
        assert g[:2] == exp[:2] and g[2] is None


def test_single_pair_dropin_api():
    """the reference's own interface (quicked_new / quicked_align / quicked_free) through the binding mirror"""
    import quicked_b200 as qb
    a = qb.QuickedAligner()
    a.align("ACGT", "ACTT")
    assert (a.getScore(), a.getCigar()) == (1, "2M1X1M")
    a.setAlgorithm(qb.BANDED); a.setBandwidth(50)
    a.align("GATTACA", "GATCACA")
    assert a.getScore() == 1
    with pytest.raises(qb.QuickedException) as e:
        a.align("", "")
    assert "Tried to align an empty sequence" in str(e.value)


def test_cigars_replay_and_ragged(gpu, oracle):
    rng = np.random.default_rng(7)
    pairs = []
    for _ in range(200):
        m = int(rng.integers(1, 600)); n = max(1, m + int(rng.integers(-30, 31)))
        pairs.append((bytes(rng.choice(list(b"ACGTN"), size=m).astype(np.uint8)),
                      bytes(rng.choice(list(b"ACGTNacgt"), size=n).astype(np.uint8))))
    for algo in (0, 1, 2, 3):
        got = gpu.align(pairs, algo=algo, bandwidth=30)
        for (p, t), g in zip(pairs, got):
            assert g == oracle.align(p, t, algo=algo, bandwidth=30)
            assert replay(expand_rle(g[2]), p.decode(), t.decode()) == g[1]


@pytest.mark.parametrize("sub_pairs", ["", "20000"])
def test_pipelined_align_batch_matches_resident_path(gpu, sub_pairs, monkeypatch):
    """qb200_align_batch cuts big batches into pipelined sub-batches (with smaller ones at both ends when there are
    enough of them: sub_pairs=20000); results must be identical and in order"""
    import ctypes as C
    import quicked_b200 as qb
    if sub_pairs:
        monkeypatch.setenv("QB200_SUB_PAIRS", sub_pairs)
    n = 230000
    seqs, po, pl, to, tl = qb.generate_pairs_native(5, n, 100, 0.05)
    gpu.upload_arrays(seqs, po, pl, to, tl)
    gpu.run(algo=0)
    status, score, off, cig = gpu.download()
    lib = qb.load()
    score2 = np.empty(n, np.int32); status2 = np.empty(n, np.int32); off2 = np.zeros(n + 1, np.int64)
    cig2 = np.zeros(int(cig.size), np.uint8)
    batch = qb.capi.Batch(seqs.ctypes.data, int(seqs.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    res = qb.capi.Results(score2.ctypes.data, status2.ctypes.data, cig2.ctypes.data, int(cig2.size), off2.ctypes.data, 0)
    p = qb.make_params(algo=0)
    assert lib.qb200_align_batch(gpu._h, C.byref(p), C.byref(batch), C.byref(res)) == 0
    assert np.array_equal(score, score2) and np.array_equal(status, status2)
    assert np.array_equal(off, off2) and res.cigar_bytes == cig.size
    assert np.array_equal(cig, cig2)


def test_cli_matches_oracle_output_format(gpu, oracle, tmp_path):
    """tools/qb_align_benchmark: same flags / .seq input / `score<TAB>CIGAR` output as the reference's align_benchmark
    (reference tools/align_benchmark/benchmark/benchmark_utils.c:151-170)."""
    import subprocess
    from quicked_b200.datagen import write_seq_file
    from _common import ROOT
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    pairs = generate_pairs(300, 500, 0.1, seed=21) + generate_pairs(20, 5000, 0.2, seed=22)
    seq = tmp_path / "in.seq"
    write_seq_file(str(seq), pairs)
    for algo, kw, extra in [("quicked", dict(algo=0), []), ("edit-banded", dict(algo=2, bandwidth=20), ["--bandwidth", "20"]),
                            ("edit-windowed", dict(algo=1), []), ("edit-banded-hirschberg", dict(algo=3, bandwidth=20), ["--bandwidth", "20"])]:
        out = tmp_path / f"{algo}.out"
        r = subprocess.run([os.path.join(ROOT, "tools", "qb_align_benchmark"), "-a", algo, "-i", str(seq), "-o", str(out), "--check", "correct",
                            "--batch-size", "128"] + extra, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = out.read_text().splitlines()
        assert len(lines) == len(pairs)
        for (p, t), line in zip(pairs, lines):
            st, sc, cg = oracle.align(p, t, **kw)
            assert line == f"{sc}\t{cg}", (algo, line[:60])


def test_cli_check_score_and_verbose(gpu, tmp_path):
    """--check score: every QUICKED score equals the exact edit distance of the tool's own multi-word checker (the
    reference asks edlib, benchmark_check.c:117-158); a 1 % band on reads with 200-long deletions is reported inexact, not failed;
    -v prints the reference's stage-timer lines (align_benchmark.c:120-129) from the GPU stage times."""
    import subprocess
    from quicked_b200.datagen import write_seq_file
    from _common import ROOT
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    pairs = generate_pairs(200, 400, 0.2, seed=41) + generate_pairs(5, 6000, 0.15, seed=42)
    seq = tmp_path / "in.seq"
    write_seq_file(str(seq), pairs)
    exe = os.path.join(ROOT, "tools", "qb_align_benchmark")
    r = subprocess.run([exe, "-a", "quicked", "-i", str(seq), "--check", "score", "-v"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert f"Score.Correct       {len(pairs)} / {len(pairs)}" in r.stderr and f"Alignments.Correct  {len(pairs)} / {len(pairs)}" in r.stderr
    assert "Time.Windowed Small" in r.stderr and "Time.Align" in r.stderr and "CIGAR.Matches" in r.stderr
    write_seq_file(str(seq), pairs + generate_pairs(30, 3000, 0.05, seed=43, indels=(4, 200)))       # 200-long deletions against a 1 % band
    r = subprocess.run([exe, "-a", "edit-banded", "--bandwidth", "1", "-i", str(seq), "--check", "alignment"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr                       # CIGARs still replay; the scores of a too-narrow band are just not optimal
    import re
    ok = int(re.search(r"Score.Correct\s+(\d+) /", r.stderr).group(1))
    assert ok <= len(pairs) + 10 and "Score.Diff" in r.stderr          # at least 20 of the 30 indel pairs are inexact


def test_cli_streaming_fasta_in_sam_out(gpu, oracle, tmp_path):
    """FASTA records (multi-line, consecutive records = pattern, text) streamed in small batches, SAM
    lines out with the reference's SAM CIGAR (query = text, reference = pattern) and NM = score; the plain
    `score<TAB>CIGAR` output of the same run must equal the .seq run's"""
    import subprocess
    from quicked_b200.capi import cigar_to_sam
    from _common import ROOT
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools")], check=True)
    pairs = [(p.decode(), t.decode()) for p, t in generate_pairs(500, 300, 0.1, seed=31) + generate_pairs(10, 4000, 0.15, seed=32)]
    pairs.append(("ACGT", ""))
    fa = tmp_path / "in.fa"
    with open(fa, "w") as f:
        for i, (p, t) in enumerate(pairs):
            for name, sq in ((f"p{i} some description", p), (f"t{i}", t)):
                f.write(f">{name}\n")
                for k in range(0, len(sq), 70):
                    f.write(sq[k:k + 70] + "\n")
    exe = os.path.join(ROOT, "tools", "qb_align_benchmark")
    for eqx in (False, True):
        out, sam = tmp_path / "fa.out", tmp_path / "fa.sam"
        r = subprocess.run([exe, "-a", "quicked", "-i", str(fa), "-o", str(out), "--output-sam", str(sam), "--batch-size", "97",
                            "--check", "correct"] + (["--sam-eqx"] if eqx else []), capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        lines = out.read_text().splitlines()
        sam_lines = [l for l in sam.read_text().splitlines() if not l.startswith("@")]
        assert len(lines) == len(pairs) and len(sam_lines) == len(pairs)
        for i, ((p, t), line, sl) in enumerate(zip(pairs, lines, sam_lines)):
            st, sc, cg = oracle.align(p, t, algo=0)
            f = sl.split("\t")
            if st < -2:                                    # empty sequence: ERROR line / unmapped SAM record
                assert line.startswith("ERROR") and f[1] == "4"
                continue
            assert line == f"{sc}\t{cg}"
            assert f[0] == f"t{i}" and f[2] == f"p{i}" and f[1] == "0" and f[3] == "1"
            assert f[5] == cigar_to_sam(cg, eqx) and f[9] == t and f[-1] == f"NM:i:{sc}"
            qlen = sum(int(n) for n, op in __import__("re").findall(r"(\d+)([MIDX=])", f[5]) if op in "MIX=")
            assert qlen == len(t)                          # SAM: query-consuming ops cover SEQ


def test_cpp_binding_example(tmp_path):
    """include/quicked.hpp: the reference's C++ binding surface (bindings/cpp/quicked.hpp:46-73) + alignMany"""
    import subprocess
    from _common import ROOT
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "tools"), "example_binding"], check=True)
    r = subprocess.run([os.path.join(ROOT, "tools", "example_binding")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout.splitlines() == ["Score: 1", "Cigar: 2M1X1M", "1\t3M1X3M", "1\t4M1D3M"]


def test_upload_device_path(gpu, oracle):
    """qb200_upload_device: the packed batch already lives in HBM (torch tensors), no host copy of the characters"""
    import torch
    import quicked_b200 as qb
    pairs = generate_pairs(300, 700, 0.1, seed=33)
    seqs, po, pl, to, tl = qb.pack_pairs(pairs)
    pad = (-len(seqs)) % 16
    seqs = np.concatenate([seqs, np.zeros(pad, np.uint8)])
    d = [torch.from_numpy(a).cuda() for a in (seqs, po, pl, to, tl)]
    torch.cuda.synchronize()
    gpu.upload_device_ptrs(d[0].data_ptr(), int(seqs.size), len(pairs), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), d[4].data_ptr())
    gpu.run(algo=0)
    status, score, off, cig = gpu.download()
    raw = cig.tobytes()
    for i, (p, t) in enumerate(pairs):
        assert (int(status[i]), int(score[i]), raw[off[i]:off[i + 1] - 1].decode()) == oracle.align(p, t)


@pytest.mark.parametrize("fused", ["0", "1"])
def test_fused_and_planned_paths_agree(oracle, fused, monkeypatch):
    """QB200_FUSED forces the single fused kernel / the three specialised kernels of the QUICKED fast path"""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", fused)
    a = qb.BatchAligner(device=0)
    pairs = generate_pairs(300, 1000, 0.1, seed=51) + generate_pairs(200, 100, 0.05, seed=52) + generate_pairs(8, 10000, 0.2, seed=53) + \
        generate_pairs(20, 3000, 0.05, seed=54, indels=(4, 200)) + [("", "ACGT"), ("ACGT", "ACGT")]
    for fs in (False, True):
        got = a.align(pairs, algo=0, force_scalar=fs)
        st = a.stats()
        assert (st["pairs_fused"] > 400) == (fused == "1")
        for (p, t), g in zip(pairs, got):
            assert g == oracle.align(p, t, algo=0, force_scalar=fs)
    a.close()


@pytest.mark.parametrize("fused", ["0", "1"])
def test_stage1_bounds_match_oracle(oracle, fused, monkeypatch):
    """WindowEd(S) score and high-error-window count of every pair (qb200_get_bounds) against the oracle's
    WindowEd(2,1): plain reads, 20 % error, long indels (walks that leave the slim quadrant slice and are redone with
    the full one), lower case / IUPAC characters on either side (raw-byte compare of equal codes)."""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", fused)
    rng = np.random.default_rng(77)
    pairs = generate_pairs(150, 1000, 0.1, seed=61) + generate_pairs(40, 3000, 0.2, seed=62) + generate_pairs(100, 100, 0.05, seed=63) + \
        generate_pairs(40, 2000, 0.05, seed=64, indels=(6, 40)) + generate_pairs(30, 700, 0.3, seed=65)
    odd = []
    for p, t in generate_pairs(60, 900, 0.08, seed=66):
        p = bytearray(p.encode() if isinstance(p, str) else p); t = bytearray(t.encode() if isinstance(t, str) else t)
        for buf in (p, t):
            for k in rng.integers(0, len(buf), size=40):
                c = chr(buf[k])
                buf[k] = ord(rng.choice(list(c.lower() + "NRYnX-")))
        odd.append((bytes(p), bytes(t)))
    pairs += odd
    a = qb.BatchAligner(device=0)
    for fs in (False, True):
        a.align(pairs, algo=0, force_scalar=fs)
        bound, hew = a.bounds()
        for i, (p, t) in enumerate(pairs):
            exp = oracle.windowed_score(p, t, 2, 1, 40, not fs)
            assert (int(bound[i]), int(hew[i])) == exp, (i, fs, len(p), len(t))
    a.close()


def test_stage1_compact_kernel(oracle, monkeypatch):
    """The COMPACT WindowEd(S) kernel (352-thread CTAs, four match-mask rows per word: the residency that makes 100 k pairs
    one wave) forced on a small batch: plain pairs go through it, pairs with an N / lower-case / IUPAC character in the
    text or an odd one in the pattern are left to the plain kernel; a pattern N stays in the compact kernel."""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", "0")
    monkeypatch.setenv("QB200_WS_COMPACT", "1")
    rng = np.random.default_rng(5)
    pairs = generate_pairs(300, 1500, 0.15, seed=91) + generate_pairs(40, 2000, 0.05, seed=92, indels=(6, 40)) + generate_pairs(60, 130, 0.1, seed=93)
    pairs = [(p.encode() if isinstance(p, str) else p, t.encode() if isinstance(t, str) else t) for p, t in pairs]
    for i in range(0, len(pairs), 7):                     # N in the pattern only
        p = bytearray(pairs[i][0]); p[int(rng.integers(0, len(p)))] = ord("N"); pairs[i] = (bytes(p), pairs[i][1])
    for i in range(3, len(pairs), 11):                    # N / odd characters in the text (the last one too: the look-ahead column)
        t = bytearray(pairs[i][1])
        for k in list(rng.integers(0, len(t), size=3)) + [len(t) - 1]:
            t[k] = ord(rng.choice(list("Nna*R")))
        pairs[i] = (pairs[i][0], bytes(t))
    a = qb.BatchAligner(device=0)
    for fs in (False, True):
        res = a.align(pairs, algo=0, force_scalar=fs)
        bound, hew = a.bounds()
        for i, (p, t) in enumerate(pairs):
            assert (int(bound[i]), int(hew[i])) == oracle.windowed_score(p, t, 2, 1, 40, not fs), (i, fs, len(p), len(t))
        for i in range(0, len(pairs), 13):
            assert res[i] == oracle.align(pairs[i][0], pairs[i][1], force_scalar=fs), i
    a.close()


@pytest.mark.parametrize("length,error,indels", [(100, 0.05, None), (1000, 0.1, None), (3000, 0.2, (5, 100)), (10000, 0.2, None), (777, 3.0, (2, 900)), (1, 0.5, None)])
def test_device_generator_is_the_host_generator(gpu, length, error, indels):
    """qb200_generate_device (one CTA per pair, edits replayed in shared memory) writes byte for byte what the host
    generator qb200_generate_pairs_ex writes for the same (seed, slice, length, error, --indels), and the batch aligns."""
    import quicked_b200 as qb
    n = 300 if length <= 3000 else 60
    for seed, first in ((7, 0), (7, 12345)):
        seqs, po, pl, to, tl = qb.generate_pairs_native(seed, n, length, error, first=first, indels=indels)
        gpu.generate_device(seed, n, length, error, first=first, indels=indels)
        dseqs, dpo, dpl, dto, dtl = gpu.download_batch()
        assert np.array_equal(dpl, pl) and np.array_equal(dtl, tl) and np.array_equal(dpo, po) and np.array_equal(dto, to)
        for i in range(n):
            assert bytes(dseqs[po[i]:po[i] + pl[i]]) == bytes(seqs[po[i]:po[i] + pl[i]]), (seed, first, i, "pattern")
            assert bytes(dseqs[to[i]:to[i] + tl[i]]) == bytes(seqs[to[i]:to[i] + tl[i]]), (seed, first, i, "text")
    gpu.run(algo=0)
    status, score, off, cig = gpu.download()
    gpu.upload_arrays(seqs, po, pl, to, tl)
    gpu.run(algo=0)
    status2, score2, off2, cig2 = gpu.download()
    assert np.array_equal(score, score2) and np.array_equal(status, status2) and np.array_equal(cig, cig2)


def test_packed_upload_matches_ascii_upload(oracle, monkeypatch):
    """2-bit packed input (qb200_pack_batch -> qb200_upload_packed): same scores and CIGARs as the ASCII upload and the
    oracle, on plain reads and on reads with N / lower-case / IUPAC characters (the exception list)."""
    import quicked_b200 as qb
    rng = np.random.default_rng(11)
    pairs = generate_pairs(120, 700, 0.12, seed=31) + generate_pairs(20, 5000, 0.2, seed=32) + generate_pairs(60, 90, 0.05, seed=33)
    pairs = [(p.encode() if isinstance(p, str) else p, t.encode() if isinstance(t, str) else t) for p, t in pairs]
    for i in range(0, len(pairs), 5):
        p, t = bytearray(pairs[i][0]), bytearray(pairs[i][1])
        for buf in (p, t):
            for k in rng.integers(0, len(buf), size=5):
                buf[k] = ord(rng.choice(list(chr(buf[k]).lower() + "NRn*")))
        pairs[i] = (bytes(p), bytes(t))
    a = qb.BatchAligner(device=0)
    for algo in (0, 2):
        got = a.align_packed(pairs, algo=algo)
        assert got == a.align(pairs, algo=algo)
        for i in range(0, len(pairs), 9):
            assert got[i] == oracle.align(pairs[i][0], pairs[i][1], algo=algo), (algo, i)
    # the pipelined form: sub-batches start on a packed byte and take their slice of the exception list
    monkeypatch.setenv("QB200_PIPELINE_MIN_PAIRS", "64")
    monkeypatch.setenv("QB200_SUB_PAIRS", "1024")
    many = generate_pairs(3000, 151, 0.1, seed=34)
    many = [(p.encode() if isinstance(p, str) else p, t.encode() if isinstance(t, str) else t) for p, t in many]
    for i in range(0, len(many), 17):
        t = bytearray(many[i][1]); t[int(rng.integers(0, len(t)))] = ord("N"); many[i] = (many[i][0], bytes(t))
    assert a.align_batch_packed(many + pairs, algo=0) == a.align(many + pairs, algo=0)
    a.close()


def test_big_batch_with_odd_characters(oracle, monkeypatch):
    """>= 16 384 pairs take the thread-per-pattern match-mask builder and (unfused) the slim WindowEd path; a tenth of
    the pairs carry lower-case / IUPAC / other bytes, which must switch those pairs to the raw-byte compare"""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", "0")
    rng = np.random.default_rng(123)
    pairs = generate_pairs(12000, 150, 0.08, seed=71) + generate_pairs(5000, 400, 0.1, seed=72)
    pairs = [(p.encode() if isinstance(p, str) else p, t.encode() if isinstance(t, str) else t) for p, t in pairs]
    for i in rng.choice(len(pairs), size=len(pairs) // 10, replace=False):
        p, t = bytearray(pairs[i][0]), bytearray(pairs[i][1])
        for buf in (p, t):
            for k in rng.integers(0, len(buf), size=6):
                buf[k] = ord(rng.choice(list(chr(buf[k]).lower() + "NRn*")))
        pairs[i] = (bytes(p), bytes(t))
    a = qb.BatchAligner(device=0)
    got = a.align(pairs, algo=0)
    bound, hew = a.bounds()
    for i, ((p, t), g) in enumerate(zip(pairs, got)):
        assert g == oracle.align(p, t, algo=0), i
    for i in rng.choice(len(pairs), size=1500, replace=False):
        assert (int(bound[i]), int(hew[i])) == oracle.windowed_score(pairs[i][0], pairs[i][1], 2, 1, 40, True), i
    a.close()


def test_small_matrix_pool_chunks_the_batch(oracle, monkeypatch):
    """qb200_set_workspace_limit below the traceback state of the batch: fill + traceback run chunk by chunk (thread
    groups, warp leaves and the host-built leaves of the slow path) and give the same answers"""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", "0")
    pairs = generate_pairs(3000, 600, 0.1, seed=81) + generate_pairs(60, 6000, 0.2, seed=82) + \
        generate_pairs(20, 3000, 0.05, seed=83, indels=(4, 200))
    a = qb.BatchAligner(device=0, workspace_limit=48 << 20)        # the batch needs ~200 MB of (Pv,Mv) entries
    for algo, kw in ((0, {}), (2, dict(bandwidth=20)), (3, dict(bandwidth=20))):
        got = a.align(pairs, algo=algo, **kw)
        for (p, t), g in zip(pairs, got):
            assert g == oracle.align(p, t, algo=algo, **kw), (algo, len(p))
    a.close()


def test_pipelined_path_with_long_reads_and_slow_pairs(oracle, monkeypatch):
    """the pipelined qb200_align_batch on a batch that also holds warp-kernel leaves, Hirschberg splits and pairs that
    go through WindowEd(L) / band doubling (forced on a small batch: 700 pairs per sub-batch)"""
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_PIPELINE_MIN_PAIRS", "1000")
    monkeypatch.setenv("QB200_SUB_PAIRS", "1024")
    pairs = generate_pairs(3000, 400, 0.1, seed=91) + generate_pairs(40, 6000, 0.2, seed=92) + \
        generate_pairs(30, 3000, 0.05, seed=93, indels=(4, 200)) + generate_pairs(3, 40000, 0.15, seed=94) + [(b"", b"ACGT")]
    rng = np.random.default_rng(3)
    order = rng.permutation(len(pairs))
    pairs = [pairs[i] for i in order]
    a = qb.BatchAligner(device=0)
    for algo, kw in ((0, {}), (3, dict(bandwidth=20))):
        got = a.align_batch(pairs, algo=algo, **kw)
        for (p, t), g in zip(pairs, got):
            assert g == oracle.align(p, t, algo=algo, **kw), (algo, len(p))
    a.close()


def test_edge_cases_match_oracle(gpu, oracle):
    """empty batch, single characters, identical / unrelated sequences, all-N, very unequal lengths, lowercase"""
    assert gpu.align([]) == []
    rng = np.random.default_rng(99)
    rnd = lambda k: bytes(rng.choice(list(b"ACGT"), size=k).astype(np.uint8))
    s500 = rnd(500)
    pairs = [("A", "A"), ("A", "C"), ("A", "ACGTACGT"), ("ACGTACGT", "A"), (s500, s500), (s500, rnd(500)), ("N" * 200, "N" * 190),
             ("N" * 100, rnd(100)), (s500.lower(), s500), (rnd(1), rnd(500)), (rnd(500), rnd(1)), (rnd(64), rnd(64)), (rnd(65), rnd(63)),
             (rnd(128), rnd(128)), (rnd(127), rnd(129)), (rnd(4096), rnd(4096)), ("ACGT" * 300, "ACGT" * 299 + "ACG"), (rnd(700), rnd(100)),
             (rnd(100), rnd(700)), ("", "A"), ("A", "")]
    for algo in (0, 1, 2, 3):
        for fs in (False, True):
            got = gpu.align(pairs, algo=algo, force_scalar=fs, bandwidth=25)
            for (p, t), g in zip(pairs, got):
                p = p.decode() if isinstance(p, bytes) else p
                t = t.decode() if isinstance(t, bytes) else t
                exp = oracle.align(p, t, algo=algo, force_scalar=fs, bandwidth=25)
                assert g == exp, (algo, fs, len(p), len(t), g[:2], exp[:2])
                if g[2]:
                    assert replay(expand_rle(g[2]), p, t) == g[1]


def test_large_batch_properties(gpu):
    """full-size style checks that do not need the oracle: scores bounded by the planted edits, CIGARs replay,
    identical pairs give the same answer wherever they sit in the batch (no cross-pair interference)"""
    import quicked_b200 as qb
    n = 60000
    seqs, po, pl, to, tl = qb.generate_pairs_native(3, n, 1000, 0.10)
    # plant duplicates of pair 0 at far-apart positions
    raw = seqs.tobytes()
    gpu.upload_arrays(seqs, po, pl, to, tl)
    gpu.run(algo=0)
    status, score, off, cig = gpu.download()
    assert (status == 1).all()
    assert (score <= 100).all() and (score >= 0).all()          # 100 planted edits are an upper bound of the distance
    text = cig.tobytes()
    for i in list(range(0, n, 997)) + [n - 1]:
        c = text[off[i]:off[i + 1] - 1].decode()
        assert replay(expand_rle(c), raw[po[i]:po[i] + pl[i]].decode(), raw[to[i]:to[i] + tl[i]].decode()) == score[i]
    # the same pairs in reverse order must give the same results
    idx = np.arange(n)[::-1].copy()
    gpu.upload_arrays(seqs, po[idx].copy(), pl[idx].copy(), to[idx].copy(), tl[idx].copy())
    gpu.run(algo=0)
    status2, score2, off2, cig2 = gpu.download()
    assert np.array_equal(score2, score[idx])
    t2 = cig2.tobytes()
    for i in range(0, n, 4999):
        j = n - 1 - i
        assert t2[off2[i]:off2[i + 1]] == text[off[j]:off[j + 1]]


def test_concurrent_aligners_from_many_threads(oracle):
    """quicked_align from several host threads at once (one aligner per thread, like the reference's OpenMP batch loop,
    align_benchmark.c:246-284): every thread has its own engine context, no global lock, results stay exact"""
    import threading
    import quicked_b200 as qb
    pairs = generate_pairs(48, 400, 0.1, seed=4242) + generate_pairs(8, 3000, 0.15, seed=4243)
    want = [oracle.align(p, t) for p, t in pairs]
    got = [None] * len(pairs)
    errors = []

    def worker(tid, nthreads):
        try:
            a = qb.QuickedAligner()
            for i in range(tid, len(pairs), nthreads):
                a.align(pairs[i][0], pairs[i][1])
                got[i] = (a.getScore(), a.getCigar())
        except Exception as e:      # pragma: no cover
            errors.append(repr(e))

    th = [threading.Thread(target=worker, args=(t, 6)) for t in range(6)]
    for t in th:
        t.start()
    for t in th:
        t.join()
    assert not errors, errors
    for g, w in zip(got, want):
        assert g == (w[1], w[2])


def test_bad_offsets_are_rejected_on_the_pipelined_path(gpu, monkeypatch):
    """a pair outside the caller's buffer is QB200_ERR_ARG on the resident AND on the pipelined path (which slices the
    buffer per sub-batch and must not read past it)"""
    import ctypes as C
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_PIPELINE_MIN_PAIRS", "64")
    monkeypatch.setenv("QB200_SUB_PAIRS", "1024")
    n = 5000
    seqs, po, pl, to, tl = qb.generate_pairs_native(6, n, 100, 0.05)
    lib = qb.load()
    score = np.empty(n, np.int32); status = np.empty(n, np.int32); off = np.zeros(n + 1, np.int64)
    cig = np.zeros(1 << 22, np.uint8)
    p = qb.make_params(algo=0)
    for bad_index, field, value in ((4321, "to", int(seqs.size) - 10), (17, "po", -5), (n - 1, "tl", -1)):
        po2, to2, tl2 = po.copy(), to.copy(), tl.copy()
        {"to": to2, "po": po2, "tl": tl2}[field][bad_index] = value
        batch = qb.capi.Batch(seqs.ctypes.data, int(seqs.size), n, po2.ctypes.data, pl.ctypes.data, to2.ctypes.data, tl2.ctypes.data)
        res = qb.capi.Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
        assert lib.qb200_align_batch(gpu._h, C.byref(p), C.byref(batch), C.byref(res)) == qb.capi.QB200_ERR_ARG
    batch = qb.capi.Batch(seqs.ctypes.data, int(seqs.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    res = qb.capi.Results(score.ctypes.data, status.ctypes.data, cig.ctypes.data, int(cig.size), off.ctypes.data, 0)
    assert lib.qb200_align_batch(gpu._h, C.byref(p), C.byref(batch), C.byref(res)) == 0


def test_empty_upload_and_workspace_smaller_than_one_group(oracle, monkeypatch):
    """qb200_upload of an empty batch (0 pairs, 0 bytes) is fine; a workspace limit below the traceback state of a single
    32-leaf thread group makes the engine grow the pool for that chunk (or fail with QB200_ERR_OOM) instead of
    writing past it"""
    import ctypes as C
    import quicked_b200 as qb
    monkeypatch.setenv("QB200_FUSED", "0")
    lib = qb.load()
    a = qb.BatchAligner(device=0, workspace_limit=1 << 20)
    empty = qb.capi.Batch(None, 0, 0, None, None, None, None)
    assert lib.qb200_upload(a._h, C.byref(empty)) == 0
    pairs = generate_pairs(200, 1000, 0.1, seed=77)          # one thread group alone: 32 x 1001 x 3 entries = 1.5 MB
    got = a.align(pairs, algo=0)
    for (p, t), g in zip(pairs, got):
        assert g == oracle.align(p, t)
    a.close()
