#!/bin/bash
# multi-GPU check under `gpurun --gpus N`: tools/gpu_multi.sh <tag> <N> [env-set ...]  (one bench run per env-set)
tag=$1; n=$2; shift 2
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_nccl.py -m gpu -x -q 2>&1 | tail -3
i=0
for v in "SX_NONE=0" "$@"; do
  i=$((i+1))
  echo "== $v"
  env $v timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29520+i)) bench.py --gpus $n --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/${tag}_$i.json 2> gpurun_out/${tag}_$i.err || tail -5 gpurun_out/${tag}_$i.err
  python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_$i.json"))
nv=d.get("nvlink") or {}
print(d["config"]["grid"], d["config"].get("exchange"), "ms/substep", round(d["ms_per_substep"],3), "Gpts/s", round(d["value"]/1e9,3), {k:round(x["ms_per_launch"],3) for k,x in d["stages"].items()}, "nvlink GB/s", round(nv.get("achieved_gbs_per_direction",0),1))
PY
done
