"""Do two/three contexts running the resident path concurrently (different streams) overlap their bottlenecks?"""
import sys, time, threading
sys.path.insert(0, ".")
import quicked_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 960000
for parts in (1, 2, 3, 4):
    gpus = [qb.BatchAligner(device=0) for _ in range(parts)]
    per = n // parts
    for k, g in enumerate(gpus):
        g.upload_arrays(*qb.generate_pairs_native(10 + k, per, 1000, 0.10))
        g.run(algo=0); g.run(algo=0)
    def work(g, reps):
        for _ in range(reps):
            g.run(algo=0)
    best = 1e9
    for rep in range(3):
        th = [threading.Thread(target=work, args=(g, 2)) for g in gpus]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        best = min(best, (time.perf_counter() - t0) / 2)
    print(f"parts={parts}: {best*1e3:.1f} ms per {n} pairs -> {n/best/1e6:.2f} M pairs/s", flush=True)
    for g in gpus: g.close()
