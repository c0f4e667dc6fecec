#!/bin/bash
# stage times of a bench workload under several env-knob settings, no test run: tools/gpu_knobs2.sh <tag> <workload> [env...]
tag=$1; wl=$2; shift 2
mkdir -p gpurun_out
for v in "SX_NONE=0" "$@"; do
  echo "== $v"
  env $v timeout 600 python bench.py --workload $wl --steps 3 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_knob.json 2>gpurun_out/${tag}_knob.err || tail -3 gpurun_out/${tag}_knob.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_knob.json"))
print(round(d["ms_per_substep"],3), {k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
PY
done
