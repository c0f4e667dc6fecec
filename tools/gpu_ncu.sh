#!/bin/bash
# one `ncu --set full` capture of part of a substep: tools/gpu_ncu.sh <tag> <skip> <count> [env...]
# kernel order in an HD substep: zinv x3, yinv x6, xpass, yfwd x3, zfwd_rk x3, project (17 launches)
tag=$1; skip=$2; count=$3; shift 3
mkdir -p gpurun_out
env "$@" timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX:-k_(zinv|yinv|xpass|yfwd|zfwd|project|zstage|inv_tma)}" -s $skip -c $count -f -o gpurun_out/${tag}_prof python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity ${BENCH_ARGS} > gpurun_out/${tag}_ncu_run.log 2>&1
tail -2 gpurun_out/${tag}_ncu_run.log; ls -la gpurun_out/${tag}_prof.ncu-rep
