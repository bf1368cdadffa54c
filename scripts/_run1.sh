mkdir -p gpurun_out/camp
python -m pytest tests -q -m gpu 2>&1 | tail -4
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py --workload c4 --steps 3 --warmup 3 > gpurun_out/camp/c4_quicked.json 2> gpurun_out/camp/c4_quicked.err
python bench.py --workload c4 --algo hirschberg --steps 3 --warmup 3 > gpurun_out/camp/c4_hirschberg.json 2> gpurun_out/camp/c4_hirschberg.err
python bench.py --workload c5 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/camp/c5_quicked.json 2> gpurun_out/camp/c5_quicked.err
python - <<PY
import json
for n in ("c4_quicked","c4_hirschberg","c5_quicked"):
    d=json.loads([l for l in open(f"gpurun_out/camp/{n}.json") if l.startswith("{")][-1])
    print(n, round(d["ms_per_step"],1), round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e_packed"] and round(d["e2e_packed"]["value"]), round(d["int_alu_roofline"]["frac"],3), d["roofline"] and round(d["roofline"]["frac"],3), d["parity"] and (d["parity"]["mismatches"], d["parity"]["cigar_errors"]), d["cpu_baseline"] and round(d["cpu_baseline"]["value"]), {k:round(v,1) for k,v in d["stage_ms_per_step"].items() if v})
PY
