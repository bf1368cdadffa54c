#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native QuickEd bound-and-align path.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation on the host cores

A "step" is one pass of the hot path (QUICKED: WindowEd bound -> BandEd/Hirschberg alignment -> score + CIGAR)
over one batch of synthetic pairs.  The default workload is the configuration BASELINE.json's target is quoted on:
configs[2], 100 k pairs of 10 kbp at 20 % error, QUICKED score + CIGAR (--algo banded / windowed give the two other
algorithms configs[2] names).  With N GPUs the SAME job is cut into contiguous index ranges, one per rank (strong
scaling; every pair has its own random stream, so a rank generates exactly its slice); there is no collective on the
data path, only the barrier + max-over-ranks of the timing.

  value    : alignments/s, whole job, inputs already resident in HBM when the timed region starts (kernel path only)
  e2e      : the same metric through qb200_align_batch() with HOST buffers — pinned-host ASCII in, host scores +
             CIGAR text out, both copies inside the timed region
  roofline : integer-ALU roofline of the dominant kernel (24 int32 ops per word-step) against the measured
             LOP3+IADD3 peak; int_alu_roofline is the same arithmetic for the whole step
  parity   : the CPU baseline leg aligns a sample of the very pairs the GPU timed; every sampled score must equal the
             reference's and a strided subset of the GPU's CIGARs is replayed against the sequences
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {   # name: (length, error, pairs of the job, description)
    "c1": (100, 0.05, 100000, "generate_dataset 100 bp pairs at 5% error, 100k pairs"),
    "c2": (1000, 0.10, 1000000, "generate_dataset 1 kbp pairs at 10% error, 1M pairs, QuickEd score + CIGAR"),
    "c3": (10000, 0.20, 100000, "generate_dataset 10 kbp pairs at 20% error, 100k pairs, score + CIGAR"),
    "c4": (100000, 0.20, 2048, "generate_dataset 100 kbp pairs at 20% error, 2048 pairs, Hirschberg CIGAR"),
    "c5": (0, 0.0, 200000, "mixed lengths 100 bp..100 kbp (equal bases per class), errors 5..25 %, length-bucketed, work-balanced ranges"),
}
C5_LENGTHS = [100, 300, 1000, 3000, 10000, 30000, 100000]
C5_ERRORS = [0.05, 0.10, 0.15, 0.20, 0.25]


def generate_c5(base_pairs, seed):
    """BASELINE config 5: one packed batch, classes in order of increasing length (length-bucketed), equal bases per
    class (pairs per class = base_pairs * 100 / L), error rate cycling 5..25 % inside every class."""
    import quicked_b200 as qb
    bufs, po, pl, to, tl, off = [], [], [], [], [], 0
    for ci, L in enumerate(C5_LENGTHS):
        per_err = max(1, base_pairs * 100 // L // len(C5_ERRORS))
        for ei, e in enumerate(C5_ERRORS):
            s_, po_, pl_, to_, tl_ = qb.generate_pairs_native(seed * 1000 + ci * 10 + ei, per_err, L, e)
            bufs.append(s_); po.append(po_ + off); to.append(to_ + off); pl.append(pl_); tl.append(tl_)
            off += s_.size
    return np.concatenate(bufs), np.concatenate(po), np.concatenate(pl), np.concatenate(to), np.concatenate(tl)
ALGOS = {"quicked": 0, "windowed": 1, "banded": 2, "hirschberg": 3}


def ncu_traffic(args):
    """{kernel name: dram__bytes_read.sum + dram__bytes_write.sum per launch} from the committed `ncu --set full` capture of
    this same command (profiles/r2_ncu_<workload>_<algo>_raw.csv); {} when the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", f"r2_ncu_{args.workload}_{args.algo}_raw.csv")
    out = {}
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        for r in rows[2:]:
            name = r[ik].split("(")[0].split("<")[0].replace("void ", "").strip().split("::")[-1]     # "void qb::k_x<1>(...)" -> "k_x"
            out[name] = out.get(name, 0.0) + float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    except Exception:
        return {}
    return out


def replay_cigar(cigar, pattern, text):
    """Replay a run-length CIGAR (M match, X mismatch, I consumes a text character, D a pattern character: reference
    cigar.c:363-434) against the two sequences.  -> its edit cost, or -1 if it does not reproduce the sequences."""
    import re
    v = h = cost = 0
    for num, op in re.findall(rb"(\d+)([MXID])", cigar):
        k = int(num)
        if op in b"MX":
            a, b = pattern[v:v + k], text[h:h + k]
            if len(a) != k or len(b) != k:
                return -1
            if op == b"M" and a != b:
                return -1
            if op == b"X" and any(x == y for x, y in zip(a, b)):
                return -1
            v += k; h += k
        elif op == b"I":
            h += k
        else:
            v += k
        if op != b"M":
            cost += k
    return cost if (v == len(pattern) and h == len(text)) else -1


def gather_pairs(seqs, po, pl, to, tl, sel):
    """the pairs `sel` of a packed batch, re-packed back to back (pattern NUL text NUL) -> (seqs, po, pl, to, tl)"""
    pl2, tl2 = np.ascontiguousarray(pl[sel]), np.ascontiguousarray(tl[sel])
    sizes = pl2.astype(np.int64) + tl2 + 2
    base = np.concatenate([[0], np.cumsum(sizes)[:-1]])
    po2 = base
    to2 = base + pl2 + 1
    total = int(sizes.sum())
    out = np.zeros((total + 15) // 16 * 16 + 16, np.uint8)
    for src, ln, dst in ((po[sel], pl2, po2), (to[sel], tl2, to2)):
        ln64 = ln.astype(np.int64)
        idx = np.repeat(src - np.concatenate([[0], np.cumsum(ln64)[:-1]]), ln64) + np.arange(int(ln64.sum()))
        jdx = np.repeat(dst - np.concatenate([[0], np.cumsum(ln64)[:-1]]), ln64) + np.arange(int(ln64.sum()))
        out[jdx] = seqs[idx]
    return out, np.ascontiguousarray(po2), pl2, np.ascontiguousarray(to2), tl2


def bind_to_gpu_numa_node(torch, local_rank):
    """One process per GPU on a multi-socket host: run this rank's threads (and so first-touch its pinned buffers) on the
    CPUs of the GPU's NUMA node, /sys/bus/pci/devices/<bdf>/local_cpulist.  The end-to-end arm streams 2 GB per step per
    GPU from host memory; across the socket interconnect the 8-GPU aggregate saturates early.  Returns the cpulist used."""
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        txt = open(f"/sys/bus/pci/devices/{bdf}/local_cpulist").read().strip()
        cpus = set()
        for part in txt.split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 4:
            os.sched_setaffinity(0, cpus)
            return txt
    except Exception:
        pass
    return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        if self.idx is None:
            return
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "250",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[1]) for r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 8 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_arm(length, error, algo_kw, sample_pairs, seed, threads=None):
    """Time the reference's CPU implementation on `sample_pairs` pairs of the workload with all host threads: ONE native
    call (oracle/_ref/libref_batch.so = the unmodified reference driven like its own align_benchmark does per thread;
    else the oracle port), so no Python overhead lands in the timed region.  -> (pairs/s, threads, kind, seconds)"""
    from oracle import harness
    seqs, po, pl, to, tl = harness.generate_pairs(seed, sample_pairs, length, error)     # oracle/datagen.c: no product code in this arm
    threads = max(1, min(threads or os.cpu_count() or 1, sample_pairs))
    harness.cpu_batch_align(seqs, po[:threads], pl[:threads], to[:threads], tl[:threads], threads, **algo_kw)   # warm the libraries
    t0 = time.perf_counter()
    kind, _, _ = harness.cpu_batch_align(seqs, po, pl, to, tl, threads, **algo_kw)
    dt = time.perf_counter() - t0
    return sample_pairs / dt, threads, kind, dt


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="strong", choices=["strong", "weak"], help="N > 1: split ONE job (strong) or one job per GPU (weak)")
    ap.add_argument("--algo", default="quicked", choices=sorted(ALGOS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs of the job (weak scaling: per GPU); default: the workload's")
    ap.add_argument("--bandwidth", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-packed", action="store_true", help="skip the 2-bit packed-input end-to-end leg")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    length, error, n_pairs, desc = WORKLOADS[args.workload]
    if args.pairs:
        n_pairs = args.pairs
    algo_kw = {"algo": ALGOS[args.algo]}
    if args.algo in ("banded", "hirschberg"):
        algo_kw["bandwidth"] = args.bandwidth
    metric, unit = "alignments_per_second", "alignments/s"
    config = {"workload": f"{args.workload}: {desc}", "algo": args.algo, "pairs_per_gpu": n_pairs, "length": length,
              "error": error, "score_and_cigar": True, "l2": "inputs_exceed_l2", "sharding": f"contiguous index ranges x{world}, no collective"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return 0
        if args.workload == "c5":
            raise SystemExit("--impl reference is defined for the uniform workloads c1..c4")
        # ~600 us/pair-thread at 10 kbp, ~50 us at 1 kbp, ~6 us at 100 bp: size each step for a few seconds
        per_pair_us = {"c1": 6, "c2": 50, "c3": 1600, "c4": 100000}[args.workload]
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or int(max(cores, min(n_pairs, 1e6 * cores / per_pair_us)))
        vals = []
        for s in range(args.warmup + args.steps):
            if s < args.warmup and s > 0:
                continue          # one warm-up pass is enough for a CPU loop
            v, used, kind, dt = cpu_reference_arm(length, error, algo_kw, sample, seed=99 + s)
            if s >= args.warmup:
                vals.append((v, dt))
        v = sum(sample for _ in vals) / sum(dt for _, dt in vals)
        ms = 1e3 * sum(dt for _, dt in vals) / len(vals)
        out = {"impl": "reference", "metric": metric, "value": v, "unit": unit, "n_gpus": args.gpus, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
               "dtype": "u64", "data": "synthetic", "config": config,
               "cpu_baseline": {"value": v, "unit": unit, "cores": used, "kind": kind,
                                "sample": f"{sample} pairs of the workload per step, {used} host threads"},
               "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "gcups_equiv": v * length * length / 1e9}
        print(json.dumps(out))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import quicked_b200 as qb
    from quicked_b200.sharding import shard_range, strided_deal
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the QuickEd GPU path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    numa = None
    if world > 1 and not os.environ.get("QB_NO_AFFINITY"):
        numa = bind_to_gpu_numa_node(torch, local_rank)       # before any pinned allocation: first touch stays node-local
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ranks(x):
        if world == 1:
            return [x]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    scaling = args.scaling if world > 1 else "strong"
    job_pairs = n_pairs
    if args.workload == "c5":
        # strong scaling of ONE mixed job: pairs sorted by estimated work and dealt round-robin to the ranks, so every GPU
        # gets the same mix of lengths and error rates (no collective; results go back by index)
        scaling = "strong"
        seqs, po, pl, to, tl = generate_c5(n_pairs, 77)
        job_pairs = int(po.size)
        sel = strided_deal(pl, tl, rank, world)
        seqs, po, pl, to, tl = gather_pairs(seqs, po, pl, to, tl, sel)
        config["job_pairs"] = job_pairs
        config["rank0_pairs"] = int(sel.size)
        config["sharding"] = f"work-sorted round-robin deal x{world}, results gathered by index, no collective"
        length = int(np.sqrt(float(np.mean(pl.astype(np.float64) * tl))))     # for the GCUPS_equiv line only
        n_pairs = int(sel.size)
    elif scaling == "strong":
        lo, hi = shard_range(job_pairs, rank, world)      # rank r aligns pairs [lo, hi) of the one job (seed 1000)
        n_pairs = hi - lo
        seqs, po, pl, to, tl = qb.generate_pairs_native(1000, n_pairs, length, error, first=lo)
        config["pairs_per_gpu"] = n_pairs
        config["job_pairs"] = job_pairs
    else:
        seqs, po, pl, to, tl = qb.generate_pairs_native(1000 + rank, n_pairs, length, error)   # one job per GPU
        job_pairs = world * n_pairs
    lib = qb.load()
    # pinned host staging for the end-to-end leg
    import ctypes as C
    pin = lib.qb200_host_alloc(seqs.size)
    pinned = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(seqs.size,))
    pinned[:] = seqs
    del seqs                          # keep one host copy per rank (8 ranks share the box's RAM)
    stream = torch.cuda.current_stream().cuda_stream
    gpu = qb.BatchAligner(device=local_rank, stream=stream)
    params = qb.make_params(**algo_kw)

    # ---- kernel path, inputs resident in HBM ----
    gpu.upload_arrays(pinned, po, pl, to, tl)
    for _ in range(args.warmup):
        gpu.run(params)
    # one nvidia-smi poller is enough (rank 0 prints the line); eight of them only compete with the ranks for the driver
    sampler = ClockSampler(None if (os.environ.get("QB_NO_SMI") or rank != 0) else local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    stage = {}
    launches = 0
    for _ in range(args.steps):
        gpu.run(params)
        st = gpu.stats()
        launches += st["kernel_launches"]
        for k, v in st.items():
            if k.startswith("ms_"):
                stage[k] = stage.get(k, 0.0) + v
    ev1.record()
    barrier()
    clocks = sampler.stop()
    my_ms = ev0.elapsed_time(ev1) / args.steps
    rank_ms = all_ranks(my_ms)
    ms_step = max(rank_ms)
    value = job_pairs / (ms_step * 1e-3)
    st = gpu.stats()
    status, score, off, cig = gpu.download()
    ok_frac = float((status >= 0).mean())
    ws_all = all_ranks(float(st["word_steps"]))

    # ---- end to end through the C-ABI with host buffers (H2D + kernels + D2H inside the timed region) ----
    score_h = np.empty(n_pairs, np.int32); status_h = np.empty(n_pairs, np.int32); off_h = np.zeros(n_pairs + 1, np.int64)
    cig_cap = int(cig.size * 1.05) + 4096 if cig is not None else 16
    cpin = lib.qb200_host_alloc(cig_cap)
    batch = qb.capi.Batch(pinned.ctypes.data, int(pinned.size), n_pairs, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data)
    res = qb.capi.Results(score_h.ctypes.data, status_h.ctypes.data, cpin, cig_cap, off_h.ctypes.data, 0)

    def e2e_step():
        rc = lib.qb200_align_batch(gpu._h, C.byref(params), C.byref(batch), C.byref(res))
        if rc != 0:
            raise RuntimeError(f"qb200_align_batch rc={rc}: {lib.qb200_last_error(gpu._h).decode()}")

    for _ in range(max(1, args.warmup)):          # the first calls create the pipeline's worker contexts and size their pools
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        e2e_step()
    torch.cuda.synchronize()
    dt = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_value = job_pairs * args.steps / dt
    st_e = gpu.stats()
    assert np.array_equal(score_h, score), "end-to-end scores differ from the resident run"
    h2d_all, d2h_all = sum(all_ranks(float(st_e["h2d_bytes"]))), sum(all_ranks(float(st_e["d2h_bytes"])))

    # ---- the same end-to-end call with the 2-bit packed input format (a quarter of the H2D bytes); the host packer runs
    #      outside the timed region and is reported beside it ----
    e2e_packed = None
    if not args.no_packed:
        t0 = time.perf_counter()
        packed, ep, ec = qb.capi.pack_2bit(pinned, po, pl, to, tl)
        pack_ms = 1e3 * (time.perf_counter() - t0)
        ppin = lib.qb200_host_alloc(packed.size)
        np.ctypeslib.as_array(C.cast(ppin, C.POINTER(C.c_uint8)), shape=(packed.size,))[:] = packed
        pbatch = qb.capi.PackedBatch(ppin, int(pinned.size), n_pairs, po.ctypes.data, pl.ctypes.data, to.ctypes.data, tl.ctypes.data,
                                     ep.ctypes.data if ep.size else None, ec.ctypes.data if ec.size else None, int(ep.size))
        score_p = np.empty(n_pairs, np.int32)
        res_p = qb.capi.Results(score_p.ctypes.data, status_h.ctypes.data, cpin, cig_cap, off_h.ctypes.data, 0)

        def packed_step():
            rc = lib.qb200_align_batch_packed(gpu._h, C.byref(params), C.byref(pbatch), C.byref(res_p))
            if rc != 0:
                raise RuntimeError(f"qb200_align_batch_packed rc={rc}: {lib.qb200_last_error(gpu._h).decode()}")

        for _ in range(max(1, args.warmup)):
            packed_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            packed_step()
        torch.cuda.synchronize()
        dtp = max_over_ranks(time.perf_counter() - t0)
        barrier()
        assert np.array_equal(score_p, score), "packed-input scores differ from the resident run"
        st_p = gpu.stats()
        e2e_packed = {"value": job_pairs * args.steps / dtp, "unit": unit,
                      "h2d_bytes_per_step": int(sum(all_ranks(float(st_p["h2d_bytes"])))),
                      "d2h_bytes_per_step": int(sum(all_ranks(float(st_p["d2h_bytes"])))),
                      "host_pack_ms_rank0": pack_ms, "exceptions_rank0": int(ep.size),
                      "note": "qb200_align_batch_packed: 2-bit stream + exception list in pinned host memory in, host scores + CIGAR text out; packing (qb200_pack_batch, all host threads) is outside the timed region"}
        lib.qb200_host_free(ppin)

    # ---- rooflines (the path is 64-bit bitwise work, bounded by the integer ALU issue rate) ----
    peaks, peak_kind = measured_peaks()
    stage_avg = {k: v / args.steps for k, v in stage.items() if k != "ms_total"}
    dom = max(stage_avg, key=stage_avg.get) if stage_avg else None
    int_peak = gpu.int_peak_tops()
    traffic = ncu_traffic(args) if (world == 1 and n_pairs == WORKLOADS[args.workload][2]) else {}
    kern = []      # (kernel, CUDA-event ms per step, word-steps per step)
    if stage_avg.get("ms_align_fill", 0) > 0:
        kern.append(("k_band_tiles (BandEd fill)" if st["word_steps_banded"] else "fill", "k_band_tiles", stage_avg["ms_align_fill"], st["word_steps_banded"]))
    if stage_avg.get("ms_windowed_s", 0) > 0:
        kern.append(("k_windowed21_score (WindowEd(S) bound)", "k_windowed21_score", stage_avg["ms_windowed_s"], st["word_steps_windowed"]))
    if stage_avg.get("ms_windowed_l", 0) > 0 and args.algo == "windowed":
        kern.append(("k_windowed_warp (WindowEd)", "k_windowed_warp", stage_avg["ms_windowed_l"], st["word_steps_windowed"]))
    if stage_avg.get("ms_fused", 0) > 0:
        kern.append(("k_quicked_fused (WindowEd(S) + fill + traceback)", "k_quicked_fused", stage_avg["ms_fused"], st["word_steps"]))
    roofs = []
    for name, key, ms, ws in kern:
        a = 24.0 * ws / (ms * 1e-3) / 1e12
        roofs.append({"kernel": name, "bound": "int_alu", "achieved": a, "peak": int_peak, "unit": "T int32-op/s",
                      "frac": a / int_peak if int_peak else None, "traffic": traffic.get(key), "ops_per_word_step": 24,
                      "word_steps_per_launch": int(ws), "ms_per_launch": ms,
                      "peak_source": "qb200_measure_int_peak (LOP3+IADD3 microbenchmark, this run); MEASURED_PEAKS.json has no integer peak"})
    for r in roofs:
        r["traffic_source"] = (f"dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this "
                               f"command, profiles/r2_ncu_{args.workload}_{args.algo}_raw.csv") if r["traffic"] is not None else None
    roofs.sort(key=lambda r: -r["ms_per_launch"])
    roof = roofs[0] if roofs else None
    ws_total = sum(ws_all)
    int_roof = {"achieved_tops": 24.0 * ws_total / (ms_step * 1e-3) / 1e12, "peak_tops": int_peak * world,
                "frac": 24.0 * ws_total / (ms_step * 1e-3) / 1e12 / (int_peak * world) if int_peak else None,
                "ops_per_word_step": 24, "word_steps_per_step": int(ws_total),
                "peak_source": "qb200_measure_int_peak (LOP3+IADD3 microbenchmark, this run)"}
    # secondary: HBM.  Algorithmic bytes of a step = the characters in + scores and CIGAR text out; the traceback state is
    # tile records (32 B per 64 word-steps), recomputed on chip by the walk
    alg_bytes = float(pinned.size) + float(cig.size if cig is not None else 0) + 8.0 * n_pairs
    hbm_roof = {"bound": "hbm", "achieved": alg_bytes / (my_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"), "unit": "GB/s",
                "frac": alg_bytes / (my_ms * 1e-3) / 1e9 / peaks.get("hbm_gbs"), "algorithmic_bytes": alg_bytes,
                "traffic": sum(traffic.values()) if traffic else None, "traffic_by_kernel": traffic or None,
                "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs"}

    # ---- CPU baseline + parity gate: the reference aligns a sample of the very pairs rank 0 timed ----
    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        from oracle import harness
        per_pair_us = {"c1": 6, "c2": 50, "c3": 1600, "c4": 100000, "c5": 400}[args.workload]
        cores = os.cpu_count() or 1
        sample = int(args.cpu_sample or max(cores, min(n_pairs, 15e6 * cores / per_pair_us)))
        sample = min(sample, n_pairs)
        harness.cpu_batch_align(pinned, po[:cores], pl[:cores], to[:cores], tl[:cores], cores, **algo_kw)      # warm the libraries
        t0 = time.perf_counter()
        kind, _, cpu_scores = harness.cpu_batch_align(pinned, po[:sample], pl[:sample], to[:sample], tl[:sample], cores, want_scores=True, **algo_kw)
        cdt = time.perf_counter() - t0
        if world == 1:
            cpu = {"value": sample / cdt, "unit": unit, "cores": min(cores, sample), "kind": kind,
                   "sample": f"the first {sample} pairs of the workload, {min(cores, sample)} host threads, {cdt:.1f} s"}
        ok = status[:sample] >= 0
        mism = int(np.count_nonzero(cpu_scores[ok] != score[:sample][ok])) + int(np.count_nonzero(~ok))
        n_replay, bad_cigars = 0, 0
        if cig is not None:
            for i in range(0, n_pairs, max(1, n_pairs // 256)):
                text = bytes(cig[off[i]:off[i + 1] - 1])
                cost = replay_cigar(text, bytes(pinned[po[i]:po[i] + pl[i]]), bytes(pinned[to[i]:to[i] + tl[i]]))
                n_replay += 1
                bad_cigars += int(cost != int(score[i]))
        parity = {"pairs_checked": sample, "mismatches": mism, "checker": kind, "cigars_replayed": n_replay, "cigar_errors": bad_cigars}

    if rank == 0:
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config,
               "gcups_equiv": value * length * length / 1e9,
               "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(h2d_all), "d2h_bytes_per_step": int(d2h_all)},
               "e2e_packed": e2e_packed, "gpu_launches": int(launches), "host_affinity": numa, "roofline": roof, "int_alu_roofline": int_roof,
               "hbm_roofline": hbm_roof, "kernel_rooflines": roofs, "cpu_baseline": cpu, "parity": parity,
               "clocks": clocks, "stage_ms_per_step": stage_avg, "dominant_stage": dom,
               "rank_ms_per_step": rank_ms, "imbalance": max(rank_ms) / (sum(rank_ms) / len(rank_ms)),
               "pairs_ok_fraction": ok_frac, "mean_score": float(score.mean()), "leaves_punted": int(st.get("leaves_punted", 0))}
        print(json.dumps(out))
    lib.qb200_host_free(pin); lib.qb200_host_free(cpin)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()
    if parity and (parity["mismatches"] or parity["cigar_errors"]):
        sys.stderr.write(f"bench.py: PARITY FAILURE {parity}\n")
        return 3
    return 0


if __name__ == "__main__":
    sys.exit(main())
