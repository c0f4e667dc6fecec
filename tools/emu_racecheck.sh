#!/bin/bash
# Race check of the kernel SOURCES without a GPU, on the CPU-thread emulation of tests/emu/cuda_emu.h:
#   SX_EMU_ADVERSARIAL bit 1 / 2: the threads of a block run in reverse / random order between barriers (a result that
#   depends on the order has a missing __syncthreads); bit 4: asynchronous copies complete as LATE as the program allows
#   (cp.async.bulk / tensor loads when a thread gets through the mbarrier wait, cp.async at wait_group, tensor-map stores
#   read their shared-memory source at wait_group.read).  mbarrier waits really wait in every mode.
# 1. the emulation suites (single rank, gloo multi-rank, C drivers) under modes 5 (reverse + late) and 6 (random + late);
# 2. self-test: two injected bugs in the bulk-copy tile kernel -- a removed mbarrier wait, a removed wait_group.read before
#    the store tile is reused -- must FAIL under mode 4 (the default emulation does not see either).
# compute-sanitizer racecheck on the B200 is the check of the real build (profiles/r2a_sanitizer_*.log, taken at the start
# of round 2); this one covers the kernels written after it.
#   tools/emu_racecheck.sh [quick]      quick: only the bulk-kernel cases in step 1
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
sel=()
[ "${1:-}" = quick ] && sel=(-k "bulk or other_fc_tables_long")
python -m specter_b200.build --emu > /dev/null
for m in 5 6; do
  echo "== emulation suites, SX_EMU_ADVERSARIAL=$m"
  SX_EMU_ADVERSARIAL=$m python -m pytest tests/test_parity_emu.py tests/test_multirank_gloo.py tests/test_c_driver.py -q -m "not gpu" "${sel[@]}" 2>&1 | tail -1
done
W=${EMU_MUT_DIR:-/tmp/specter_emu_mut}
mutant() {   # name, sed expression on sx_fused_tiles.cu
  rm -rf "$W/$1" && mkdir -p "$W/$1" && cp -r specter_b200 include "$W/$1/" && mkdir -p "$W/$1/tests" && cp -r tests/emu "$W/$1/tests/"
  sed -i "$2" "$W/$1/specter_b200/csrc/sx_fused_tiles.cu"
  if cmp -s "$W/$1/specter_b200/csrc/sx_fused_tiles.cu" specter_b200/csrc/sx_fused_tiles.cu; then echo "mutant $1: the pattern no longer matches"; return 1; fi
  objs=""
  for src in "$W/$1"/specter_b200/csrc/*.cu; do
    o="$W/$1/$(basename "${src%.cu}").o"
    g++ -O2 -std=c++20 -fPIC -DSX_EMU -U_FORTIFY_SOURCE -D_FORTIFY_SOURCE=0 -include "$W/$1/tests/emu/cuda_emu.h" -x c++ -c "$src" -o "$o" -Wno-unknown-pragmas &
    objs="$objs $o"
  done
  wait
  g++ -shared -Wl,-Bsymbolic -o "$W/$1/lib.so" $objs -lpthread -latomic
  for m in 0 4; do
    r=$(SPECTER_EMU_LIB="$W/$1/lib.so" SX_EMU_ADVERSARIAL=$m python -m pytest tests/test_parity_emu.py -q -x -k hd_substeps_bulk_tiles 2>&1 | tail -1)
    echo "mutant $1, SX_EMU_ADVERSARIAL=$m: $r"
  done
}
echo "== self-test: injected bugs (expected: passed under 0, FAILED under 4)"
mutant no_mbarrier_wait 's/^    mbar_wait(bar0 + cs, (phase >> cs) \& 1u);$/    ;/'
mutant no_wait_group_read 's/make_hook(\[&\] { if (lead) tma_store_wait_read(); }/make_hook([\&] { }/'
