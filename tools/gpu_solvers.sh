#!/bin/bash
# BOUSS / MHD single-GPU stage times at 512^3
mkdir -p gpurun_out
for w in bouss512 mhd512; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 > gpurun_out/${1:-s}_$w.json 2> gpurun_out/${1:-s}_$w.err || tail -5 gpurun_out/${1:-s}_$w.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${1:-s}_$w.json"))
print("$w", round(d["ms_per_substep"],3), round(d["roofline"]["frac"],3), {k:(round(v["ms_per_launch"],3), v["launches_per_substep"]) for k,v in d["stages"].items()})
PY
done
