#!/bin/bash
# Round-2 call 20 (1 GPU): one 24-warp march CTA per SM (93 KB of shared memory instead of 3 x 48 KB -> more L1) and a
# single A tile in the bucketed feature kernel (145 KB instead of 193 KB).
mkdir -p gpurun_out
T=c20
. tools/run_fn.sh
echo "== GPU tests"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
run SNRF_X=0
run SNRF_MARCH_WPC=8
run "SNRF_LIB_PATH=$PWD/libsnrf_a2.so"
run SNRF_MARCH_CARVEOUT=44
run SNRF_MARCH_CARVEOUT=58
run SNRF_X=0 --chunk 640000
