"""Dev probe: one QUICKED run of a synthetic batch with the tile kernels' timing counters (QB200_TILE_DEBUG)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("QB200_TILE_DEBUG", "1")
import numpy as np
import quicked_b200 as qb

length = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
err = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
n = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
algo = sys.argv[4] if len(sys.argv) > 4 else "quicked"
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 3
seqs, po, pl, to, tl = qb.generate_pairs_native(1234, n, length, err)
a = qb.BatchAligner(device=0)
kw = {"algo": {"quicked": 0, "windowed": 1, "banded": 2, "hirschberg": 3}[algo]}
if algo in ("banded", "hirschberg"):
    kw["bandwidth"] = 20
a.upload_arrays(seqs, po, pl, to, tl)
for r in range(reps):
    a.run(**kw)
    st = a.stats()
    print({k: (round(v, 3) if isinstance(v, float) else v) for k, v in st.items() if k.startswith("ms_") or k in ("word_steps", "leaves", "leaves_punted", "kernel_launches")})
