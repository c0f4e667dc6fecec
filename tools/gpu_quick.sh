#!/bin/bash
# quick perf check: parity tests + stage times of the 512^3 bench (no e2e / cpu legs), optional env knobs
tag=${1:-q}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
run() {
  echo "== $*"
  env "$@" timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_knob.json 2>gpurun_out/${tag}_knob.err || tail -3 gpurun_out/${tag}_knob.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_knob.json"))
print(round(d["ms_per_substep"],3), {k:round(v["ms_per_launch"],3) for k,v in d["stages"].items()})
PY
}
run SX_NONE=0
for v in "$@"; do run $v; done
