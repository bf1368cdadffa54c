python -m pytest tests -x -q -m gpu 2>&1 | tail -5
for cfg in "c3 100000" "c3 12500" "c2 1000000"; do set -- $cfg; python bench.py --workload $1 --pairs $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2i_$1_$2.json 2> gpurun_out/r2i_$1_$2.err; tail -c 300 gpurun_out/r2i_$1_$2.err; done
python - <<PY
import json
for n in ("c3_100000","c3_12500","c2_1000000"):
    d=json.loads([l for l in open(f"gpurun_out/r2i_{n}.json") if l.startswith("{")][-1])
    print(n, round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), "packed", d["e2e_packed"] and (round(d["e2e_packed"]["value"]), d["e2e_packed"]["h2d_bytes_per_step"], round(d["e2e_packed"]["host_pack_ms_rank0"])))
PY
