#!/bin/bash
# three small `ncu --set full` captures of one substep (report sizes stay under the gpurun 64 MiB limit):
#   a: first zinv + first two yinv (derivative / plain)   b: xpass + first yfwd   c: last zfwd_rk + project
tag=$1; shift
EXTRA=("$@")
mkdir -p gpurun_out
cap() {  # name skip count
  env "${EXTRA[@]}" SX_DUMMY=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_(zinv|yinv|inv_tma|xpass|yfwd|zfwd|project)' -s $2 -c $3 -f -o gpurun_out/${tag}_$1 python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_$1.log 2>&1
  ls -la gpurun_out/${tag}_$1.ncu-rep
}
# kernel order in an HD substep: zinv x3, yinv x6, xpass, yfwd x3, zfwd_rk x3, project (17 launches)
cap a 17 5
cap b 26 2
cap c 32 2
