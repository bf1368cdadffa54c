python -m pytest tests -x -q -m gpu 2>&1 | tail -3
for cfg in "c3 12500" "c4 2048" "c3 100000"; do set -- $cfg; python bench.py --workload $1 --pairs $2 --steps 4 --warmup 2 --no-cpu-baseline --no-packed > gpurun_out/r2v_$1_$2.json 2> gpurun_out/r2v_$1_$2.err; tail -c 200 gpurun_out/r2v_$1_$2.err; done
python - <<PY
import json
for n in ("c3_12500","c4_2048","c3_100000"):
    d=json.loads([l for l in open(f"gpurun_out/r2v_{n}.json") if l.startswith("{")][-1])
    print(n, round(d["ms_per_step"],2), round(d["value"]), {k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v}, round(d["int_alu_roofline"]["frac"],3))
PY
