#!/usr/bin/env python
"""oracle/gen_golden.py — golden vectors dumped from the UNMODIFIED reference (oracle/_ref/libquicked_ref.so).

Run in the build container (where /root/reference exists):   python oracle/gen_golden.py
Writes tests/golden/golden_explicit.json (sequences inline) and tests/golden/golden_seeded.json
(sequences regenerated from quicked_b200.datagen seeds; outputs stored as score + sha1(cigar)).
Inputs on which the reference itself is undefined (oracle flag `ref_undefined`) are left out.
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.harness import Oracle, Reference, build  # noqa: E402
from quicked_b200.datagen import generate_pairs, read_seq_file  # noqa: E402

ALGOS = {"quicked": 0, "windowed": 1, "banded": 2, "hirschberg": 3}
SEEDED = [  # (name, num, length, error, seed, indels, {algo: params})
    ("c1_100bp_5pct", 200, 100, 0.05, 101, None),
    ("c2_1kbp_10pct", 100, 1000, 0.10, 102, None),
    ("c3_10kbp_20pct", 16, 10000, 0.20, 103, None),
    ("c4_100kbp_20pct", 2, 100000, 0.20, 104, None),
    ("indels_3kbp", 40, 3000, 0.05, 105, (4, 200)),
    ("indels_10kbp", 12, 10000, 0.10, 106, (4, 400)),
    ("mixed_300bp_25pct", 60, 300, 0.25, 107, None),
    ("mixed_30kbp_15pct", 3, 30000, 0.15, 108, None),
]
PARAMS = {"quicked": {}, "banded": {"bandwidth": 20}, "windowed": {}, "hirschberg": {"bandwidth": 20},
          "windowed_2_1": {"algo": 1, "window_size": 2, "overlap_size": 1},
          "banded_5": {"algo": 2, "bandwidth": 5}}


def sha(s):
    return hashlib.sha1((s or "").encode()).hexdigest()[:16]


def main():
    build(ref=True)
    o, r = Oracle(), Reference()
    gold_dir = os.path.join(ROOT, "tests", "golden")
    os.makedirs(gold_dir, exist_ok=True)

    explicit = []
    fixed = [("GATC", "GATO"), ("ACGT", "ACTT"), ("A", "A"), ("A", "C"), ("ACGTACGTAC", "ACGT"), ("ACGT", "ACGTACGTAC"),
             ("N" * 70, "ACGT" * 20), ("acgtnACGTN" * 13, "ACGTNacgtn" * 12)]
    fixed += generate_pairs(24, 64, 0.1, seed=1) + generate_pairs(24, 130, 0.2, seed=2) + generate_pairs(12, 257, 0.08, seed=3)
    for p, t in fixed:
        p = p.decode() if isinstance(p, bytes) else p
        t = t.decode() if isinstance(t, bytes) else t
        rec = {"pattern": p, "text": t, "out": {}}
        for name, prm in PARAMS.items():
            kw = dict(prm)
            kw.setdefault("algo", ALGOS.get(name, 0))
            if o.align(p, t, full=True, **kw)[3]["ref_undefined"]:
                continue
            st, sc, cg = r.align(p, t, **kw)
            rec["out"][name] = {"status": st, "score": sc, "cigar": cg}
        explicit.append(rec)
    json.dump({"source": "oracle/_ref/libquicked_ref.so via oracle/gen_golden.py", "params": PARAMS, "cases": explicit},
              open(os.path.join(gold_dir, "golden_explicit.json"), "w"), indent=0)

    seeded = []
    for name, num, length, err, seed, indels in SEEDED:
        pairs = generate_pairs(num, length, err, seed=seed, indels=indels)
        rec = {"name": name, "num": num, "length": length, "error": err, "seed": seed, "indels": indels,
               "inputs_sha1": sha("".join(p.decode() + "|" + t.decode() + "\n" for p, t in pairs)), "out": {}}
        for aname, prm in PARAMS.items():
            kw = dict(prm)
            kw.setdefault("algo", ALGOS.get(aname, 0))
            rows = []
            for p, t in pairs:
                if o.align(p, t, full=True, **kw)[3]["ref_undefined"]:
                    rows.append(None)
                    continue
                st, sc, cg = r.align(p, t, **kw)
                rows.append([st, sc, sha(cg)])
            rec["out"][aname] = rows
        seeded.append(rec)
        print("golden", name, "done", flush=True)
    ont = os.path.join(gold_dir, "ONT.MiniION.1.seq")
    ont_rec = None
    if os.path.exists(ont):
        p, t = read_seq_file(ont)[0]
        st, sc, cg = r.align(p, t)
        ont_rec = {"status": st, "score": sc, "cigar_sha1": sha(cg), "m": len(p), "n": len(t)}
    json.dump({"source": "oracle/_ref/libquicked_ref.so via oracle/gen_golden.py", "params": PARAMS,
               "sets": seeded, "ont": ont_rec}, open(os.path.join(gold_dir, "golden_seeded.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
