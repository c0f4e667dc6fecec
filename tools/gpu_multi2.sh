#!/bin/bash
# multi-GPU bench lines under `gpurun --gpus N`: tools/gpu_multi2.sh <tag> <N> "<bench args>" [more "<bench args>" ...]
# (bench args may start with ENV=VALUE words); the NCCL parity test first when asked with TESTS=1
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader | head -2
free -g | head -2
if [ "${TESTS:-0}" = "1" ]; then timeout 900 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3; fi
i=0
for args in "$@"; do
  i=$((i+1))
  envs=""; rest=""
  for w in $args; do case $w in [A-Z_]*=*) envs="$envs $w";; *) rest="$rest $w";; esac; done
  echo "== [$i] $args"
  env $envs timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $n $rest > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err || tail -8 gpurun_out/${tag}_$i.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_$i.json"))
    nv=d.get("nvlink") or {}
    print(d["config"]["name"], d["config"]["grid"], d["config"].get("exchange"), "ms/substep", round(d["ms_per_substep"],3), "Gpts/s", round(d["value"]/1e9,3),
          "whole", round(d["roofline"]["whole_substep"]["frac"],3), {k:round(x["ms_per_launch"],3) for k,x in d["stages"].items()},
          "nvlink GB/s", round(nv.get("achieved_gbs_per_direction",0),1), "parity", (d.get("parity_check") or {}).get("ok"), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
except Exception as e:
    print("no line:", e)
PY
done
