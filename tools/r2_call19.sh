#!/bin/bash
# Round-2 call 19 (1 GPU): L1 policy of the march kernel's gathers (no-allocate brick loads from level L on, no-allocate
# table loads of the un-bricked hashed levels) and the shared-memory carve-out.
mkdir -p gpurun_out
. tools/run_fn.sh
run SNRF_X=0
for L in libsnrf_s8.so libsnrf_s10.so libsnrf_s12.so libsnrf_s10h.so libsnrf_h.so; do
  run "SNRF_LIB_PATH=$PWD/$L"
done
run SNRF_MARCH_CARVEOUT=58
run SNRF_MARCH_CARVEOUT=72
run SNRF_MARCH_CARVEOUT=100
