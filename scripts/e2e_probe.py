"""Probe the pipelined end-to-end path: python scripts/e2e_probe.py [n_pairs]   (env QB200_WORKERS, QB200_TRACE)"""
import ctypes as C, sys, time
import numpy as np
sys.path.insert(0, ".")
import quicked_b200 as qb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
lib = qb.load()
seqs, po, pl, to, tl = qb.generate_pairs_native(1, n, 1000, 0.10)
pin = lib.qb200_host_alloc(seqs.size); pinned = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(seqs.size,)); pinned[:] = seqs
gpu = qb.BatchAligner(device=0)
params = qb.make_params(algo=0)
cap = 500 * n
cpin = lib.qb200_host_alloc(cap)
score = np.empty(n, np.int32); status = np.empty(n, np.int32); off = np.zeros(n + 1, np.int64)
batch = qb.capi.Batch(pinned.ctypes.data, int(pinned.size), n, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
res = qb.capi.Results(score.ctypes.data, status.ctypes.data, cpin, cap, off.ctypes.data, 0)
for it in range(4):
    t0 = time.perf_counter()
    rc = lib.qb200_align_batch(gpu._h, C.byref(params), C.byref(batch), C.byref(res))
    dt = time.perf_counter() - t0
    print(f"iter {it}: rc={rc} {dt*1e3:.1f} ms -> {n/dt/1e6:.2f} M pairs/s, cigar bytes {res.cigar_bytes}", flush=True)
