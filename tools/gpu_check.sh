#!/bin/bash
# Run on the GPU box through gpurun:  tests, bench, ncu launch list, one ncu --set full capture.
# Usage: tools/gpu_check.sh <tag> [what...]   what in: tests bench modular launches ncu
tag=${1:-r1}; shift
what=${*:-tests bench launches ncu}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.max.mem,power.limit --format=csv > gpurun_out/${tag}_gpu.txt 2>&1
for w in $what; do
  case $w in
    tests) timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_pytest_gpu.log 2>&1; tail -5 gpurun_out/${tag}_pytest_gpu.log;;
    smoke) timeout 600 python __graft_entry__.py smoke > gpurun_out/${tag}_smoke.log 2>&1; tail -2 gpurun_out/${tag}_smoke.log;;
    bench) timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; cat gpurun_out/${tag}_bench.json; tail -3 gpurun_out/${tag}_bench.err;;
    modular) timeout 900 python bench.py --path 1 --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_bench_modular.json 2> gpurun_out/${tag}_bench_modular.err; cat gpurun_out/${tag}_bench_modular.json;;
    launches) timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_launches_run.log 2>&1; tail -2 gpurun_out/${tag}_launches_run.log;;
    ncu) timeout 1200 ncu --set full --clock-control none --import-source on -k regex:'k_(zinv|yinv|xpass|yfwd|zfwd|project)' -s 17 -c 17 -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu_run.log 2>&1; tail -2 gpurun_out/${tag}_ncu_run.log; ls -la gpurun_out/${tag}_prof.ncu-rep;;
    reference) timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>&1; cat gpurun_out/${tag}_bench_reference.json;;
  esac
done
