"""Quick per-stage timing of the resident (kernel-only) path on synthetic configs.  Usage:
   python scripts/quick_perf.py [c1|c2|c3|...] [n_pairs]"""
import sys
import time

sys.path.insert(0, ".")
import quicked_b200 as qb  # noqa: E402

CONFIGS = {"c1": (100, 0.05, 100000), "c2": (1000, 0.10, 200000), "c3": (10000, 0.20, 10000), "c2s": (1000, 0.10, 20000), "c4": (100000, 0.20, 256)}


def main():
    names = [a for a in sys.argv[1:] if a in CONFIGS] or ["c1", "c2", "c3"]
    nums = [int(a) for a in sys.argv[1:] if a.isdigit()]
    algo = 0
    for a in sys.argv[1:]:
        if a.startswith("algo="):
            algo = int(a[5:])
    gpu = qb.BatchAligner(device=0)
    for name in names:
        length, err, n = CONFIGS[name]
        if nums:
            n = nums[0]
        t0 = time.time()
        arrays = qb.generate_pairs_native(1234, n, length, err)
        t1 = time.time()
        gpu.upload_arrays(*arrays)
        t2 = time.time()
        for it in range(3):
            gpu.run(algo=algo, bandwidth=20)
            st = gpu.stats()
            print(name, "run", it, {k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if v})
        t3 = time.time()
        status, score, off, cig = gpu.download()
        t4 = time.time()
        ms = st["ms_total"]
        print(f"{name}: n={n} gen {t1-t0:.2f}s upload {t2-t1:.2f}s download {t4-t3:.2f}s | kernel-path {ms:.2f} ms -> "
              f"{n/ms*1e3:.0f} pairs/s, {st['cells']/ms/1e6:.1f} GCUPS_equiv, {st['word_steps']/ms/1e6:.2f} G word-steps/s "
              f"| mean score {score.mean():.1f} status ok {(status>=0).mean():.3f}")


if __name__ == "__main__":
    main()
