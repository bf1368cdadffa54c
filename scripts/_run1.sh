for v in "QB200_SUB_PAIRS=33334 QB200_COMPUTE_THREADS=2" "QB200_SUB_PAIRS=12500 QB200_COMPUTE_THREADS=2" "QB200_SUB_PAIRS=12500 QB200_COMPUTE_THREADS=3" "QB200_SUB_PAIRS=16667 QB200_COMPUTE_THREADS=3" "QB200_SUB_PAIRS=25000 QB200_COMPUTE_THREADS=3" "QB200_SUB_PAIRS=8192 QB200_COMPUTE_THREADS=3 QB200_WORKERS=6"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-packed > gpurun_out/r2o.json 2> gpurun_out/r2o.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open("gpurun_out/r2o.json") if l.startswith("{")][-1])
    print("$v", "e2e", round(d["e2e"]["value"]), "resident", round(d["value"]))
except Exception as e:
    print("$v", "ERR", open("gpurun_out/r2o.err").read()[-300:])
PY
done
