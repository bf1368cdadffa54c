mkdir -p gpurun_out/camp
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 8 --workload c4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/camp/c4_quicked_8gpu.json 2> gpurun_out/camp/c4_quicked_8gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/camp/c4_quicked_8gpu.json") if l.startswith("{")][-1])
print("c4 8gpu ms", round(d["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), d["e2e_packed"] and round(d["e2e_packed"]["value"]), "imb", round(d["imbalance"],3), {k:round(v,1) for k,v in d["stage_ms_per_step"].items() if v})
PY
