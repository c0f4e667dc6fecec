#!/bin/bash
# ncu --set full captures of the round's closing state, in two calls (gpurun brings back at most 64 MiB per call):
#   tools/gpu_r2p.sh <tag> hd   : bench line + HD 512^3: zinv / yinv (a), zfwd_rk (c)
#   tools/gpu_r2p.sh <tag> hd2  : HD 512^3: xpass / yfwd (b), project (d), with the source pages
#   tools/gpu_r2p.sh <tag> mhd  : MHD cross-product x passes, vector-potential boundary kernel; BOUSS four-component x pass
tag=${1:-r2p}; part=${2:-hd}
mkdir -p gpurun_out
cap() {  # name workload skip count [extra ncu args]
  local name=$1 wl=$2 skip=$3 count=$4; shift 4
  timeout 900 ncu --set full --clock-control none "$@" -k regex:'k_(zinv|yinv|inv_tma|xpass|yfwd|zfwd|project|aproject)' -s $skip -c $count -f -o gpurun_out/${tag}_$name python bench.py --workload $wl --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-parity > gpurun_out/${tag}_$name.log 2>&1
  ls -la gpurun_out/${tag}_$name.ncu-rep
}
if [ $part = hd ]; then
  timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err || tail -3 gpurun_out/${tag}_bench.err
  # kernel order in an HD substep: zinv x3, yinv x6, xpass, yfwd x3, zfwd_rk x3, project (17 launches)
  cap a hd512 17 5
  cap c hd512 32 1
elif [ $part = hd2 ]; then
  cap b hd512 26 2 --import-source on
  cap d hd512 33 1 --import-source on
else
  # MHD substep: zinv x12, yinv x12, xcross x2, yfwd x6, zfwd_rk x3, project, zfwd_rk x3, aproject (40 launches)
  cap mhd_x mhd512 64 2
  cap mhd_p mhd512 79 1 --import-source on
  # BOUSS substep: zinv x4, yinv x8, xpass, yfwd x4, zfwd_rk x4, project (22 launches)
  cap bouss_x bouss512 34 1
fi
# the raw metric pages as CSV next to the reports; reports are dropped, largest first, if the call would exceed what
# gpurun brings back (64 MiB)
for r in gpurun_out/${tag}_*.ncu-rep; do ncu -i $r --page raw --csv > ${r%.ncu-rep}_raw.csv 2>/dev/null; done
while [ $(du -sm gpurun_out | cut -f1) -ge 60 ]; do
  big=$(ls -S gpurun_out/*.ncu-rep 2>/dev/null | head -1); [ -z "$big" ] && break
  echo "dropping $big"; rm -f $big
done
du -sh gpurun_out
