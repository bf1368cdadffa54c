"""Two host threads, each with its own context, calling the blocking qb200_align_batch on alternating batches: does the
fill / drain of one call overlap with the steady state of the other?   python scripts/e2e_probe2.py [n_pairs] [threads]"""
import ctypes as C, sys, time, threading
import numpy as np
sys.path.insert(0, ".")
import quicked_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
T = int(sys.argv[2]) if len(sys.argv) > 2 else 2
lib = qb.load()
seqs, po, pl, to, tl = qb.generate_pairs_native(1, n, 1000, 0.10)
params = qb.make_params(algo=0)
workers = []
pin = lib.qb200_host_alloc(seqs.size)
assert pin, "pinned input allocation failed"
pinned = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(seqs.size,)); pinned[:] = seqs   # shared, read-only
for t in range(T):
    gpu = qb.BatchAligner(device=0)
    cap = 420 * n
    cpin = lib.qb200_host_alloc(cap)
    assert cpin, "pinned output allocation failed"
    score = np.empty(n, np.int32); status = np.empty(n, np.int32); off = np.zeros(n + 1, np.int64)
    batch = qb.capi.Batch(pinned.ctypes.data, int(pinned.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    res = qb.capi.Results(score.ctypes.data, status.ctypes.data, cpin, cap, off.ctypes.data, 0)
    workers.append((gpu, batch, res, score, status, off, cpin))
def run(w, reps):
    gpu, batch, res = w[0], w[1], w[2]
    for _ in range(reps):
        rc = lib.qb200_align_batch(gpu._h, C.byref(params), C.byref(batch), C.byref(res))
        assert rc == 0
for w in workers: run(w, 2)          # warm-up (allocations)
for rep in range(3):
    K = 4
    th = [threading.Thread(target=run, args=(w, K)) for w in workers]
    t0 = time.perf_counter()
    for t in th: t.start()
    for t in th: t.join()
    dt = time.perf_counter() - t0
    print(f"threads={T}: {K*T} batches of {n} in {dt*1e3:.1f} ms -> {dt*1e3/(K*T):.1f} ms per batch, {K*T*n/dt/1e6:.2f} M pairs/s", flush=True)
