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
for v in "$@"; do run $v; done
