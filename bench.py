#!/usr/bin/env python
"""bench.py — headline benchmark of the B200-native QuickEd bound-and-align path.

  python bench.py --gpus N --steps K --warmup W            our arm (one rank per GPU; torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...   the reference's own CPU implementation on the host cores

A "step" is one pass of the hot path (QUICKED: WindowEd bound -> BandEd/Hirschberg alignment -> score + CIGAR)
over one batch of synthetic pairs.  The default workload is BASELINE.json configs[1]: 1 kbp pairs at 10 % error,
1 M pairs per GPU (weak scaling: every rank aligns its own contiguous index range; there is no collective on the
data path, only the barrier + max-over-ranks of the timing).

  value : alignments/s, whole job, inputs already resident in HBM when the timed region starts (kernel path only)
  e2e   : the same metric through qb200_align_batch() with HOST buffers — pinned-host ASCII in, host scores +
          CIGAR text out, both copies inside the timed region
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

WORKLOADS = {   # name: (length, error, pairs per GPU, description)
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


def ncu_traffic_bytes(kernel_substr, csv_name="r1_ncu_full_c2_v7_raw.csv"):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel_substr`, from the committed `ncu --set full`
    capture of this same command (profiles/, 1 M pairs of configs[1]); None when the capture is not there."""
    import csv
    path = os.path.join(ROOT, "profiles", csv_name)
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ik, ir, iw = hdr.index("Kernel Name"), hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
        for r in rows[2:]:
            if kernel_substr in r[ik]:
                return float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]]
    except Exception:
        return None
    return None


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
    import quicked_b200 as qb
    seqs, po, pl, to, tl = qb.generate_pairs_native(seed, sample_pairs, length, error)
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
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--algo", default="quicked", choices=sorted(ALGOS))
    ap.add_argument("--pairs", type=int, default=0, help="pairs per GPU (default: the workload's)")
    ap.add_argument("--bandwidth", type=int, default=20)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto, ~10-30 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
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
               "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
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

    scaling = "weak"
    if args.workload == "c5":
        # strong scaling: ONE mixed job, length-bucketed, cut into contiguous work-balanced ranges (no collective)
        from quicked_b200.sharding import balanced_ranges
        seqs, po, pl, to, tl = generate_c5(n_pairs, 77)
        lo, hi = balanced_ranges(list(zip(pl.tolist(), tl.tolist())), world)[rank]
        b0 = int(min(po[lo], to[lo])) & ~15
        b1 = int(max(po[hi - 1] + pl[hi - 1], to[hi - 1] + tl[hi - 1]))
        seqs = np.ascontiguousarray(np.concatenate([seqs[b0:b1], np.zeros((-(b1 - b0)) % 16 + 16, np.uint8)]))
        po, to, pl, tl = po[lo:hi] - b0, to[lo:hi] - b0, np.ascontiguousarray(pl[lo:hi]), np.ascontiguousarray(tl[lo:hi])
        config["job_pairs"] = int(sum(max(1, n_pairs * 100 // L // len(C5_ERRORS)) * len(C5_ERRORS) for L in C5_LENGTHS))
        config["rank0_pairs"] = int(hi - lo)
        length = int(np.sqrt(float(np.mean(pl.astype(np.float64) * tl))))     # for the GCUPS_equiv line only
        n_pairs = int(hi - lo)
        scaling = "strong"
    else:
        # rank r aligns the contiguous index range [r*n_pairs, (r+1)*n_pairs) of the (virtual) job
        seqs, po, pl, to, tl = qb.generate_pairs_native(1000 + rank, n_pairs, length, error)
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
    # one sampler per rank on its own GPU; QB_NO_SMI=1 disables it (nvidia-smi polling can perturb short steps)
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
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    ms_step = ms_total / args.steps
    total_pairs = world * n_pairs
    if scaling == "strong" and world > 1:
        tp = torch.tensor([float(n_pairs)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tp)
        total_pairs = int(tp.item())
    elif scaling == "strong":
        total_pairs = n_pairs
    value = total_pairs / (ms_step * 1e-3)
    st = gpu.stats()
    status, score, off, cig = gpu.download()
    ok_frac = float((status >= 0).mean())

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
    e2e_value = total_pairs * args.steps / dt
    st_e = gpu.stats()
    assert np.array_equal(score_h, score), "end-to-end scores differ from the resident run"

    # ---- roofline of the dominant kernel ----
    peaks, peak_kind = measured_peaks()
    stage_avg = {k: v / args.steps for k, v in stage.items() if k != "ms_total"}
    dom = max(stage_avg, key=stage_avg.get) if stage_avg else None
    int_peak = gpu.int_peak_tops()
    # the traceback-state fill: every word-step writes its 16-byte (Pv,Mv) entry (SURVEY §8d "spilled-matrix": 16 B/word-step)
    fill_ms = stage_avg.get("ms_align_fill", 0.0)
    fill_bytes = 16.0 * st["word_steps_banded"]
    roof = None
    if fill_ms > 0:
        ach = fill_bytes / (fill_ms * 1e-3) / 1e9
        roof = {"kernel": "k_banded_thread/k_banded_warp (BandEd full-matrix fill)", "bound": "hbm", "achieved": ach,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "frac": ach / peaks.get("hbm_gbs"),
                "traffic": ncu_traffic_bytes("k_banded_thread") if (args.workload == "c2" and n_pairs == 1000000 and args.algo == "quicked") else None,
                "algorithmic_bytes": fill_bytes,
                "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs", "algorithmic_bytes_per_word_step": 16,
                "ms_per_launch": fill_ms}
    int_roof = {"achieved_tops": 24.0 * st["word_steps"] / (ms_step * 1e-3) / 1e12, "peak_tops": int_peak,
                "frac": 24.0 * st["word_steps"] / (ms_step * 1e-3) / 1e12 / int_peak if int_peak else None,
                "ops_per_word_step": 24, "word_steps_per_step": st["word_steps"],
                "peak_source": "qb200_measure_int_peak (LOP3+IADD3 microbenchmark, this run)"}

    # the other two big kernels, for the record (not the `roofline` contract object): WindowEd(S) against the measured
    # integer peak, the traceback against HBM (16-byte entry per visited text column; it is latency-, not bandwidth-bound)
    others = []
    if stage_avg.get("ms_windowed_s", 0) > 0 and int_peak:
        a = 24.0 * st["word_steps_windowed"] / (stage_avg["ms_windowed_s"] * 1e-3) / 1e12
        others.append({"kernel": "k_windowed21_score (WindowEd(S) bound)", "bound": "int_alu", "achieved": a, "peak": int_peak,
                       "unit": "T int32-op/s", "frac": a / int_peak, "ms_per_launch": stage_avg["ms_windowed_s"]})
    if stage_avg.get("ms_align_trace", 0) > 0 and args.workload in ("c1", "c2"):
        tb = 16.0 * n_pairs * length
        a = tb / (stage_avg["ms_align_trace"] * 1e-3) / 1e9
        others.append({"kernel": "k_traceback_thread (BandEd traceback)", "bound": "hbm_latency", "achieved": a, "peak": peaks.get("hbm_gbs"),
                       "unit": "GB/s", "frac": a / peaks.get("hbm_gbs"), "algorithmic_bytes": tb, "ms_per_launch": stage_avg["ms_align_trace"],
                       "traffic": ncu_traffic_bytes("k_traceback_thread") if (args.workload == "c2" and n_pairs == 1000000 and args.algo == "quicked") else None})

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and args.workload != "c5":
        per_pair_us = {"c1": 6, "c2": 50, "c3": 1600, "c4": 100000, "c5": 50}[args.workload]
        cores = os.cpu_count() or 1
        sample = args.cpu_sample or int(max(cores, min(n_pairs, 15e6 * cores / per_pair_us)))
        v, used, kind, cdt = cpu_reference_arm(length, error, algo_kw, sample, seed=1000)
        cpu = {"value": v, "unit": unit, "cores": used, "kind": kind,
               "sample": f"{sample} pairs of the workload, {used} host threads, {cdt:.1f} s"}

    if rank == 0:
        out = {"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
               "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "u64",
               "data": "synthetic", "config": config,
               "gcups_equiv": value * length * length / 1e9,
               "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": int(st_e["h2d_bytes"]),
                       "d2h_bytes_per_step": int(st_e["d2h_bytes"])},
               "gpu_launches": int(launches), "host_affinity": numa, "roofline": roof, "int_alu_roofline": int_roof, "other_kernel_rooflines": others, "cpu_baseline": cpu,
               "clocks": clocks, "stage_ms_per_step": stage_avg, "dominant_stage": dom,
               "pairs_ok_fraction": ok_frac, "mean_score": float(score.mean())}
        print(json.dumps(out))
    lib.qb200_host_free(pin); lib.qb200_host_free(cpin)
    gpu.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
