"""Multi-GPU sharding of a batch: contiguous index ranges per rank, no collective on the data path.

The reference parallelises only across independent pairs (OpenMP `parallel for` over a batch,
reference tools/align_benchmark/align_benchmark.c:269-284); the B200 equivalent is one process per GPU, each
aligning its own contiguous slice, followed by a host-side gather of (score, CIGAR) back into input order."""


def shard_range(n, rank, world):
    """[lo, hi) of rank's contiguous slice; sizes differ by at most one"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_results(scores, cigars, rank, world):
    """gather per-shard python lists on rank 0 in input order (torch.distributed object gather; gloo or nccl)"""
    if world == 1:
        return list(scores), list(cigars)
    import torch.distributed as dist
    out = [None] * world if rank == 0 else None
    dist.gather_object((list(scores), list(cigars)), out, dst=0)
    if rank != 0:
        return None, None
    all_s, all_c = [], []
    for s, c in out:
        all_s += s
        all_c += c
    return all_s, all_c


def align_sharded(pairs, rank, world, device=None, **params):
    """Align rank's slice of `pairs` on its GPU and gather everything on rank 0.  -> (status, score, cigar) list or None"""
    from .capi import BatchAligner
    lo, hi = shard_range(len(pairs), rank, world)
    gpu = BatchAligner(device=rank if device is None else device)
    res = gpu.align(pairs[lo:hi], **params)
    gpu.close()
    sc, cg = gather_results([(r[0], r[1]) for r in res], [r[2] for r in res], rank, world)
    if rank != 0:
        return None
    return [(s[0], s[1], c) for s, c in zip(sc, cg)]


def estimated_work(m, n, error=0.15):
    """Rough word-step count of QUICKED for one pair: WindowEd(S) (2 words x 2 passes over the text) plus a band of
    ceil(error * max(m, n) / 64) + 2 blocks over n columns (twice when Hirschberg splits).  Only the relative values
    matter: it is the key for balancing mixed-length batches across GPUs (BASELINE config 5)."""
    L = max(m, n)
    band = -(-int(error * L) // 64) + 2
    split = 2 if band * n * 16 > (1 << 24) else 1
    return 4 * n + band * n * split


def balanced_ranges(lengths, world, error=0.15):
    """Contiguous index ranges [lo, hi) per rank with approximately equal estimated work (prefix-sum split).
    `lengths` = [(m, n), ...] in input order; contiguous ranges keep the host-side gather a plain concatenation."""
    work = [estimated_work(m, n, error) for m, n in lengths]
    total = float(sum(work)) or 1.0
    bounds, acc, nxt = [0], 0.0, 1
    for i, w in enumerate(work):
        acc += w
        while nxt < world and acc >= total * nxt / world:
            bounds.append(i + 1)
            nxt += 1
    while len(bounds) < world:
        bounds.append(len(work))
    bounds.append(len(work))
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def strided_deal(pattern_len, text_len, rank, world, error=0.15):
    """Indices of rank's share of a mixed-length batch: pairs sorted by estimated work (heaviest first) and dealt
    round-robin, so every GPU gets the same mix of lengths (BASELINE config 5); results are gathered by index.
    numpy arrays in, sorted numpy index array out."""
    import numpy as np
    m = np.asarray(pattern_len, dtype=np.int64); n = np.asarray(text_len, dtype=np.int64)
    L = np.maximum(m, n)
    band = -(-(error * L).astype(np.int64) // 64) + 2
    split = np.where(band * n * 16 > (1 << 24), 2, 1)
    work = 4 * n + band * n * split
    order = np.argsort(-work, kind="stable")
    return np.sort(order[rank::world])
