#!/bin/bash
mkdir -p gpurun_out
echo "== full"; SX_XP=2 SX_PJ=1 timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/knob.json 2>gpurun_out/knob.err
python -c "
import json; d=json.load(open('gpurun_out/knob.json')); print(round(d['ms_per_substep'],3), {k:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
echo "== no FFT (data movement only)"; SPECTER_B200_LIB=$PWD/gpurun_nofft.so timeout 600 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/knob_nofft.json 2>gpurun_out/knob.err
python -c "
import json; d=json.load(open('gpurun_out/knob_nofft.json')); print(round(d['ms_per_substep'],3), {k:round(v['ms_per_launch'],3) for k,v in d['stages'].items()})"
