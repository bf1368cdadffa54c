"""Seeded twin of the reference's `generate_dataset` edit model (tools/generate_dataset/generate_dataset.c).

The reference tool is unseeded (`srand(time(0))`, generate_dataset.c:375); this module reproduces its
*model* deterministically so tests and benchmarks can regenerate the same pairs anywhere:

  * text    = `length` uniform ACGT characters (generate_dataset.c:52-63);
  * pattern = a copy of the text with `ceil(length * error)` edits applied one after another, each chosen
              uniformly among mismatch / deletion / insertion at a uniform position of the current string
              (generate_dataset.c:108-199); `error >= 1` is an absolute count (generate_dataset.c:370);
  * optional `indels=(num, length)` large insertions/deletions (the `--indels` switch).

Output convention of the tool: '>' line = pattern, '<' line = text (generate_dataset.c:396-405).
"""
import math
import numpy as np

ALPHABET = np.frombuffer(b"ACGT", dtype=np.uint8)


def _mutate(seq, num_errors, rng):
    seq = bytearray(seq)
    for _ in range(num_errors):
        kind = int(rng.integers(0, 3))
        n = len(seq)
        if kind == 0 and n > 0:      # mismatch: redraw until the base changes
            pos = int(rng.integers(0, n))
            while True:
                ch = int(ALPHABET[rng.integers(0, 4)])
                if ch != seq[pos]:
                    break
            seq[pos] = ch
        elif kind == 1 and n > 1:    # deletion
            pos = int(rng.integers(0, n))
            del seq[pos]
        else:                        # insertion of a random base
            pos = int(rng.integers(0, max(n, 1)))
            seq.insert(pos, int(ALPHABET[rng.integers(0, 4)]))
    return seq


def _large_indels(seq, num, length, rng):
    seq = bytearray(seq)
    for _ in range(num):
        if rng.integers(0, 2) == 0 and len(seq) > length + 1:
            pos = int(rng.integers(0, len(seq) - length))
            del seq[pos:pos + length]
        else:
            pos = int(rng.integers(0, len(seq) + 1))
            seq[pos:pos] = ALPHABET[rng.integers(0, 4, size=length)].tobytes()
    return seq


def generate_pair(length, error, rng, indels=None):
    """-> (pattern: bytes, text: bytes)"""
    text = ALPHABET[rng.integers(0, 4, size=length)].tobytes()
    num_errors = int(error) if error >= 1.0 else int(math.ceil(length * error))
    pattern = _mutate(text, num_errors, rng)
    if indels:
        pattern = _large_indels(pattern, indels[0], indels[1], rng)
    if len(pattern) == 0:
        pattern = bytearray(b"A")
    return bytes(pattern), text


def generate_pairs(num, length, error, seed=0, indels=None):
    rng = np.random.default_rng(seed)
    return [generate_pair(length, error, rng, indels) for _ in range(num)]


def write_seq_file(path, pairs):
    """The `.seq` format read by align_benchmark (align_benchmark.c:73-99)."""
    with open(path, "wb") as f:
        for p, t in pairs:
            f.write(b">" + p + b"\n<" + t + b"\n")


def read_seq_file(path, limit=None):
    pairs = []
    with open(path, "rb") as f:
        while True:
            l1 = f.readline()
            l2 = f.readline()
            if not l1 or not l2:
                break
            pairs.append((l1[1:].rstrip(b"\n"), l2[1:].rstrip(b"\n")))
            if limit and len(pairs) >= limit:
                break
    return pairs
