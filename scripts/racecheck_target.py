"""small racecheck target: warp text kernel, tile traceback, WindowEd(S) compact, tile fill, device generator"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["QB200_FUSED"] = "0"
os.environ["QB200_WS_COMPACT"] = "1"
import quicked_b200 as qb
g = qb.BatchAligner(device=0)
g.generate_device(3, 48, 5000, 0.15)
for algo in (0, 1):
    g.run(algo=algo)
    st, sc, off, cig = g.download()
    print(algo, int(sc.sum()), int(cig.size))
g.close()
