#!/bin/bash
# Round-2 call 15 (1 GPU): split march (sampling half at 5 CTAs/SM + field half), 4 CTAs/SM variants, dynamic tile
# scheduler of the bucketed feature kernel, one launch per frame.
mkdir -p gpurun_out
T=c15
. tools/run_fn.sh
echo "== GPU tests (split on)"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; tail -3 gpurun_out/${T}_tests.log
B="--brick-gb 62"
run SNRF_MARCH_SPLIT=0 $B
run SNRF_MARCH_SPLIT=1 $B
for L in libsnrf_m4.so libsnrf_m4p6.so libsnrf_p4.so; do
  echo "-- $L"
  run "SNRF_LIB_PATH=$PWD/$L SNRF_MARCH_SPLIT=0" $B
  run "SNRF_LIB_PATH=$PWD/$L SNRF_MARCH_SPLIT=1" $B
done
echo "-- one launch per frame"
run SNRF_MARCH_SPLIT=1 $B --chunk 640000
run SNRF_MARCH_SPLIT=0 $B --chunk 640000
echo "== tests under the 4-CTA library"; SNRF_LIB_PATH=$PWD/libsnrf_m4.so timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests_m4.log 2>&1; tail -3 gpurun_out/${T}_tests_m4.log
