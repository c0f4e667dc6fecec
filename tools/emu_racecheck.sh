#!/bin/bash
# Race check of the kernel SOURCES without a GPU, on the CPU-thread emulation of tests/emu/cuda_emu.h:
#   SX_EMU_ADVERSARIAL bit 1 / 2: the threads of a block run in reverse / random order between barriers (a result that
#   depends on the order has a missing __syncthreads); bit 4: asynchronous copies complete as LATE as the program allows
#   (cp.async.bulk / tensor loads when a thread gets through the mbarrier wait, cp.async at wait_group, tensor-map stores
#   read their shared-memory source at wait_group.read); bit 8: STREAMS are queues that run as late and as little as the
#   program allows -- an operation executes only when the host waits for it or when something the host waits for depends
#   on it through cudaStreamWaitEvent, so work that no event orders before its consumer has not run when the consumer
#   does (pageable / page-locked host memory follow the CUDA rules); bit 16 (with 8): the communication and copy streams
#   of the library run as EARLY as their event waits allow while the compute stream stays lazy -- a buffer overwritten
#   before its readers are done changes the result.  mbarrier waits really wait in every mode.
# 1. the emulation suites (single rank, gloo multi-rank incl. the peer-to-peer pipeline, C drivers, ABI) under modes 13
#    (reverse + late copies + lazy streams), 14 (random + late copies + lazy streams) and 30 (14 + eager side streams);
# 2. self-test: injected bugs must FAIL under the adversary and are invisible to the default emulation -- in the bulk-copy
#    tile kernel a removed mbarrier wait and a removed wait_group.read before the store tile is reused (mode 4); on the
#    host side a removed wait of the compute stream for the upload event of sx_hd_step_host and a removed wait for the
#    completion event of the peer-to-peer exchange (mode 8), and the removed wait of sx_hd_step_host for the work in flight
#    before it overwrites the state (mode 24).
# compute-sanitizer racecheck on the B200 is the check of the real build (profiles/r2a_sanitizer_*.log, taken at the start
# of round 2); this one covers the kernels written after it.
#   tools/emu_racecheck.sh [quick]      quick: only the bulk-kernel cases in step 1
set -u
ROOT=$(cd "$(dirname "$0")/.." && pwd)
cd "$ROOT"
sel=()
[ "${1:-}" = quick ] && sel=(-k "bulk or other_fc_tables_long or p2p or step_host")
python -m specter_b200.build --emu > /dev/null
for m in 13 14 30; do
  echo "== emulation suites, SX_EMU_ADVERSARIAL=$m"
  SX_EMU_ADVERSARIAL=$m python -m pytest tests/test_parity_emu.py tests/test_multirank_gloo.py tests/test_c_driver.py tests/test_abi.py -q -m "not gpu" "${sel[@]}" 2>&1 | tail -1
done
W=${EMU_MUT_DIR:-/tmp/specter_emu_mut}
mutant() {   # name, file under specter_b200/csrc, sed expression, adversarial mode that must catch it, test file, -k expression
  rm -rf "$W/$1" && mkdir -p "$W/$1" && cp -r specter_b200 include "$W/$1/" && mkdir -p "$W/$1/tests" && cp -r tests/emu "$W/$1/tests/"
  sed -i "$3" "$W/$1/specter_b200/csrc/$2"
  if cmp -s "$W/$1/specter_b200/csrc/$2" "specter_b200/csrc/$2"; then echo "mutant $1: the pattern no longer matches"; return 1; fi
  objs=""
  for src in "$W/$1"/specter_b200/csrc/*.cu; do
    o="$W/$1/$(basename "${src%.cu}").o"
    g++ -O2 -std=c++20 -fPIC -DSX_EMU -U_FORTIFY_SOURCE -D_FORTIFY_SOURCE=0 -include "$W/$1/tests/emu/cuda_emu.h" -x c++ -c "$src" -o "$o" -Wno-unknown-pragmas &
    objs="$objs $o"
  done
  wait
  g++ -shared -Wl,-Bsymbolic -o "$W/$1/lib.so" $objs -lpthread -latomic
  for m in 0 "$4"; do
    r=$(SPECTER_EMU_LIB="$W/$1/lib.so" SX_EMU_ADVERSARIAL=$m python -m pytest "$5" -q -x -k "$6" 2>&1 | tail -1)
    echo "mutant $1, SX_EMU_ADVERSARIAL=$m: $r"
  done
}
echo "== self-test: injected bugs (expected: passed under 0, FAILED under the adversary)"
mutant no_mbarrier_wait sx_fused_tiles.cu 's/^    mbar_wait(bar0 + cs, (phase >> cs) \& 1u);$/    ;/' 4 tests/test_parity_emu.py hd_substeps_bulk_tiles
mutant no_wait_group_read sx_fused_tiles.cu 's/make_hook(\[&\] { if (lead) tma_store_wait_read(); }/make_hook([\&] { }/' 4 tests/test_parity_emu.py hd_substeps_bulk_tiles
mutant no_upload_event_wait sx_fused.cu 's|^    SX_CUDA_CHECK(cudaStreamWaitEvent(p.stream, p.pre_wait\[i\], 0));$|    ;|' 8 tests/test_parity_emu.py hd_step_host
mutant no_exchange_done_wait sx_comm.cu 's|^  SX_CUDA_CHECK(cudaStreamWaitEvent(p.stream, c.done\[ev\], 0));.*$|  ;|' 8 tests/test_multirank_gloo.py "substep_multirank_p2p and 2-variants0"
mutant no_sync_before_overwrite sx_rkstep.cu 's|^  SX_CUDA_CHECK(cudaStreamSynchronize(p.stream));   // nothing in flight still reads the state that is overwritten$|  ;|' 24 tests/test_parity_emu.py hd_step_host
