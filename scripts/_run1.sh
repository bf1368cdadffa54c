for p in 75000 100000 150000; do
python bench.py --workload c3 --algo windowed --pairs $p --steps 4 --warmup 2 --no-cpu-baseline --no-packed > gpurun_out/r2q.json 2> gpurun_out/r2q.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2q.json") if l.startswith("{")][-1])
print($p, round(d["ms_per_step"],1), {k:round(v,1) for k,v in d["stage_ms_per_step"].items() if v})
PY
done
