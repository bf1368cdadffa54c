N=$1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_c3_${N}gpu.json 2> gpurun_out/r2g_c3_${N}gpu.err
tail -c 400 gpurun_out/r2g_c3_${N}gpu.err
python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2g_c3_${N}gpu.json") if l.startswith("{")][-1])
print($N, "ms", round(d["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "imb", round(d["imbalance"],3), d["rank_ms_per_step"], {k:round(v,2) for k,v in d["stage_ms_per_step"].items()})
PY
