#!/bin/bash
# tuning experiment: kernel variants on the 512^3 bench (no e2e / cpu legs)
mkdir -p gpurun_out
run() {
  echo "== $*"
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/knob.json 2>gpurun_out/knob.err || tail -3 gpurun_out/knob.err
  python - <<PY
import json
d=json.load(open("gpurun_out/knob.json"))
print(round(d["ms_per_substep"],3), {k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
PY
}
run SX_XP=0 SX_PJ=0 SX_ZF=0
run SX_XP=1 SX_PJ=1 SX_ZF=1
run SX_XP=2 SX_PJ=2
run SX_XP=3 SX_PJ=3
run SX_XP=4 SX_PJ=4
