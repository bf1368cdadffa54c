for v in "A" "QB200_TILE_OCC=5" "QB200_TILE_OCC=3" "QB200_TILE_LANES=256" "QB200_TILE_LANES=256 QB200_TILE_SLOTS=32" "QB200_TILE_OCC=5 QB200_TILE_SLOTS=16"; do
  env $v python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-packed > gpurun_out/r2m.json 2> gpurun_out/r2m.err
  python - <<PY
import json
d=json.loads([l for l in open("gpurun_out/r2m.json") if l.startswith("{")][-1])
print("$v", round(d["ms_per_step"],2), {k:round(v,2) for k,v in d["stage_ms_per_step"].items() if v})
PY
done
