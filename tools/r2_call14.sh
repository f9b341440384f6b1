#!/bin/bash
# Round-2 call 14 (1 GPU): march kernel with the field MLPs on tcgen05 (SNRF_MARCH_TC), FHADD lerps, bitonic top-k,
# CTA-aggregated bucket pre-pass, deeper brick budgets.  A/B against the previous library (libsnrf_base.so).
mkdir -p gpurun_out
T=c14
. tools/run_fn.sh
echo "== GPU tests (TC on)"; timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests.log 2>&1; tail -5 gpurun_out/${T}_tests.log
if ! grep -q " passed" gpurun_out/${T}_tests.log || grep -q " failed" gpurun_out/${T}_tests.log; then
  echo "== GPU tests (TC off)"; SNRF_MARCH_TC=0 timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/${T}_tests_notc.log 2>&1; tail -5 gpurun_out/${T}_tests_notc.log
fi
run SNRF_LIB_PATH=$PWD/libsnrf_base.so
run SNRF_MARCH_TC=0
run SNRF_MARCH_TC=1
run SNRF_MARCH_TC=0 --brick-gb 24
run SNRF_MARCH_TC=1 --brick-gb 24
run SNRF_MARCH_TC=0 --brick-gb 62
run SNRF_MARCH_TC=1 --brick-gb 62
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
