#!/bin/bash
# tuning experiment: tile-kernel variants on the 512^3 bench (no e2e / cpu legs)
mkdir -p gpurun_out
for cfg in "8 1" "8 2" "4 2" "4 4"; do
  set -- $cfg
  echo "== NP=$1 MINB=$2"
  SX_TILE_NP=$1 SX_TILE_MINB=$2 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/knob_$1_$2.json 2>gpurun_out/knob_$1_$2.err
  python - <<PY
import json
d=json.load(open("gpurun_out/knob_$1_$2.json"))
print(d["ms_per_substep"], {k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
PY
done
