#!/bin/bash
# Round-2 call 21 (1 GPU): wide march CTA with a contiguous ray range per CTA and a shared-memory ray counter.
mkdir -p gpurun_out
T=c21
. tools/run_fn.sh
echo "== GPU tests"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
run SNRF_X=0
run SNRF_MARCH_WPC=8
run SNRF_X=0 --chunk 640000
run SNRF_MARCH_WPC=8 --chunk 640000
