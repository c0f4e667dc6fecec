#!/bin/bash
# experiment build: tools/build_variant.sh <name> <-DDEFINE ...>  ->  gpurun_<name>.so at the repo root
# (A/B timing of build-time switches such as -DSX_NOFFT or -DSX_TW_TABLE; never loaded by default)
set -e
name=$1; shift
root="$(cd "$(dirname "$0")/.." && pwd)"
cd "$root/specter_b200/csrc"
mkdir -p /tmp/sxvar_$name
for f in *.cu; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr "$@" -Xcompiler -fPIC -c $f -o /tmp/sxvar_$name/${f%.cu}.o 2>&1 | grep -v deprecated &
done
wait
nvcc -shared -Xlinker -Bsymbolic -o "$root/gpurun_$name.so" /tmp/sxvar_$name/*.o -lcudart -lnccl -ldl 2>&1 | grep -v deprecated
ls -la "$root/gpurun_$name.so"
