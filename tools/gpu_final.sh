#!/bin/bash
# closing check of a round on one B200: GPU suite, smoke, the bench line (both arms)
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${tag}_pytest_gpu.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -1 gpurun_out/${tag}_smoke.log
timeout 600 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err; cut -c1-300 gpurun_out/${tag}_bench_reference.json
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -3 gpurun_out/${tag}_bench.err
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("hd512 ms/substep", round(d["ms_per_substep"],3), "whole", round(d["roofline"]["whole_substep"]["frac"],3), "dominant", d["roofline"]["kernel"], round(d["roofline"]["frac"],3), "traffic", d["roofline"]["traffic"], d["roofline"]["traffic_from_profile"], "e2e", d["e2e"]["ms_per_step"], "clocks", d["clocks"], "cpu", d["cpu_baseline"]["value"], "state_check", d.get("state_check"))
PY
