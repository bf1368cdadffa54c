python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 400 compute-sanitizer --tool racecheck --error-exitcode 9 python scripts/_race.py > gpurun_out/r2_racecheck.txt 2>&1; echo "exit $?" >> gpurun_out/r2_racecheck.txt
tail -6 gpurun_out/r2_racecheck.txt
for cfg in "c3 100000" "c3 12500" "c2 1000000"; do set -- $cfg; python bench.py --workload $1 --pairs $2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2n_$1_$2.json 2> gpurun_out/r2n_$1_$2.err; tail -c 300 gpurun_out/r2n_$1_$2.err; done
python - <<PY
import json
for n in ("c3_100000","c3_12500","c2_1000000"):
    d=json.loads([l for l in open(f"gpurun_out/r2n_{n}.json") if l.startswith("{")][-1])
    print(n, round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), {k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v}, round(d["int_alu_roofline"]["frac"],3))
PY
