#!/bin/bash
# full GPU suite + the three single-GPU bench lines + ncu launch lists per solver
tag=${1:-r2l}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -4 gpurun_out/${tag}_pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -3 gpurun_out/${tag}_bench.err
for w in bouss512 mhd512; do
  timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --no-parity > gpurun_out/${tag}_$w.json 2> gpurun_out/${tag}_$w.err || tail -5 gpurun_out/${tag}_$w.err
done
python - <<PY
import json
for w in ("bench","bouss512","mhd512"):
    d=json.load(open("gpurun_out/${tag}_%s.json"%w))
    print(w, "ms/substep", round(d["ms_per_substep"],3), "whole", round(d["roofline"]["whole_substep"]["frac"],3), "dominant", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "e2e", (d.get("e2e") or {}).get("ms_per_step"))
    print("   ", {k:(round(v["ms_per_launch"],3), v["launches_per_substep"]) for k,v in d["stages"].items()})
PY
for w in hd512 bouss512 mhd512; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/${tag}_launches_$w.csv python bench.py --workload $w --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_launches_$w.log 2>&1; tail -1 gpurun_out/${tag}_launches_$w.log | cut -c1-200
done
