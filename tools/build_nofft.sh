#!/bin/bash
# experiment build: same kernels with the FFT butterflies compiled out (data movement only)
set -e
cd "$(dirname "$0")/../specter_b200/csrc"
mkdir -p /tmp/nofft
for f in sx_api sx_kernels_fft sx_kernels_ops sx_rkstep sx_fused sx_comm; do
  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo --expt-relaxed-constexpr -DSX_NOFFT -Xcompiler -fPIC -c $f.cu -o /tmp/nofft/$f.o &
done
wait
nvcc -shared -Xlinker -Bsymbolic -o ../../gpurun_nofft.so /tmp/nofft/*.o -lcudart -lnccl -ldl
