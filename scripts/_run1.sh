mkdir -p gpurun_out/camp
python bench.py --steps 10 --warmup 3 > gpurun_out/camp/c3_quicked.json 2> gpurun_out/camp/c3_quicked.err
tail -c 300 gpurun_out/camp/c3_quicked.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/camp/c3_quicked.json") if l.startswith("{")][-1])
print(round(d["ms_per_step"],2), round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e_packed"]["value"]), round(d["int_alu_roofline"]["frac"],4), round(d["roofline"]["frac"],3), d["parity"], d["cpu_baseline"]["value"], {k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v})
PY
