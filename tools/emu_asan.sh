#!/bin/bash
# Host-side memory check without a GPU: the kernel SOURCES compiled for the CPU-thread emulation (tests/emu) with
# AddressSanitizer, and the emulation parity suite run on that build.  Device allocations and the dynamic shared memory
# of every emulated block are heap blocks there, so an out-of-bounds index in a kernel or a launcher is reported.
# (compute-sanitizer on the B200 stays the check of the real build: profiles/r2a_sanitizer_*.log.)
#   tools/emu_asan.sh [pytest args]        default: tests/test_parity_emu.py
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
OUT=${EMU_ASAN_DIR:-/tmp/specter_emu_asan}
mkdir -p "$OUT"
objs=""
for src in "$ROOT"/specter_b200/csrc/*.cu; do
  o="$OUT/$(basename "${src%.cu}").o"
  g++ -O1 -g -fsanitize=address -fno-omit-frame-pointer -std=c++20 -fPIC -DSX_EMU -DSX_EMU_UCONTEXT_ONLY -include "$ROOT/tests/emu/cuda_emu.h" \
      -x c++ -c "$src" -o "$o" -Wno-unknown-pragmas &
  objs="$objs $o"
done
wait
g++ -shared -fsanitize=address -Wl,-Bsymbolic -o "$OUT/libspecter_emu_asan.so" $objs -lpthread -latomic
cd "$ROOT"
LD_PRELOAD=$(gcc -print-file-name=libasan.so) ASAN_OPTIONS=detect_leaks=0:detect_stack_use_after_return=0 \
SPECTER_EMU_LIB="$OUT/libspecter_emu_asan.so" python -m pytest -q -x -m "not gpu" "${@:-tests/test_parity_emu.py}"
