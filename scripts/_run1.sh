python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2l_c3.json 2> gpurun_out/r2l_c3.err; tail -c 300 gpurun_out/r2l_c3.err
python - <<PY
import json
d=json.loads([l for l in open(f"gpurun_out/r2l_c3.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
PY
