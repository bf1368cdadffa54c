# multi-GPU lines of the round: bash scripts/campaign_ngpu.sh N [all]
N=$1
O=gpurun_out/camp
mkdir -p $O
run() { # name, bench args
  name=$1; shift
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > $O/${name}_${N}gpu.json 2> $O/${name}_${N}gpu.err
  grep -c '^{' $O/${name}_${N}gpu.json
}
run c3_quicked --steps 10 --warmup 3 --no-cpu-baseline
if [ "$2" = "all" ]; then
  run c4_quicked --workload c4 --steps 3 --warmup 3 --no-cpu-baseline
  run c5_quicked --workload c5 --steps 3 --warmup 3 --no-cpu-baseline
  run c3_quicked_weak --scaling weak --steps 5 --warmup 3 --no-cpu-baseline
  run c3_windowed --algo windowed --steps 5 --warmup 3 --no-cpu-baseline
  run c3_banded --algo banded --steps 5 --warmup 3 --no-cpu-baseline
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$O/*_${N}gpu.json")):
    try:
        d=json.loads([l for l in open(f) if l.startswith("{")][-1])
        print(f.split("/")[-1], "ms", round(d["ms_per_step"],2), "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "packed", d.get("e2e_packed") and round(d["e2e_packed"]["value"]), "imb", round(d["imbalance"],3), d["scaling"])
    except Exception as e:
        print(f, "ERR", e)
PY
