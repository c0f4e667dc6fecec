#!/bin/bash
# Round-2 opening call on one B200: toolchain probe of the GPU box, GPU parity suite, the three single-GPU BASELINE
# workloads (hd512 full line, bouss512, mhd512), compute-sanitizer on the small cases.
# Usage: tools/gpu_r2a.sh <tag> [what...]   what in: probe tests hd bouss mhd sanitize
tag=${1:-r2a}; shift
what=${*:-probe tests hd bouss mhd sanitize}
mkdir -p gpurun_out
for w in $what; do
  case $w in
    probe)
      { echo "== toolchain"; which gfortran mpif90 mpifort mpicc mpirun flang ifort nvfortran 2>&1; ldconfig -p | grep -i -E "fftw|libmpi" ;
        ls /usr/lib/x86_64-linux-gnu | grep -i -E "fftw|openmpi|mpich" ; echo "== host"; nproc; free -g | head -2; lscpu | grep -E "Model name|Socket|Thread|Core" ;
        echo "== gpu"; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit,memory.total --format=csv; nvidia-smi topo -m 2>&1 | head -20; } > gpurun_out/${tag}_probe.txt 2>&1
      cat gpurun_out/${tag}_probe.txt;;
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log;;
    hd) timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err;;
    bouss|mhd)
      timeout 900 python bench.py --workload ${w}512 --steps 3 --warmup 3 > gpurun_out/${tag}_${w}512.json 2> gpurun_out/${tag}_${w}512.err || tail -5 gpurun_out/${tag}_${w}512.err
      python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_${w}512.json"))
print("${w}512", "ms/substep", round(d["ms_per_substep"],3), "whole-substep frac", round(d["roofline"]["whole_substep"]["frac"],3), "parity", d.get("parity_check"))
for k,v in d["stages"].items(): print("   ", k, round(v["ms_per_launch"],3), "x", v["launches_per_substep"], "frac", round(v.get("frac_of_hbm_peak",0),3))
PY
      ;;
    sanitize)
      for tool in memcheck racecheck; do
        SX_TMA_MIN=64 timeout 600 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_${tool}_smoke.log python __graft_entry__.py smoke > gpurun_out/${tag}_sanitizer_${tool}_smoke.out 2>&1
        tail -3 gpurun_out/${tag}_sanitizer_${tool}_smoke.log
        timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/${tag}_sanitizer_${tool}_256.log python tools/sanitize_case.py > gpurun_out/${tag}_sanitizer_${tool}_256.out 2>&1
        tail -3 gpurun_out/${tag}_sanitizer_${tool}_256.log; tail -2 gpurun_out/${tag}_sanitizer_${tool}_256.out
      done;;
  esac
done
