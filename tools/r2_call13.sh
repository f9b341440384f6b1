#!/bin/bash
# Round-2 call 13 (1 GPU): nn.Module shims on hardware, block-of-4 hashed gather variant A/B, brick budget 10 vs 4.
mkdir -p gpurun_out
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c13_tests.log 2>&1; tail -5 gpurun_out/c13_tests.log
. tools/run_fn.sh
run SNRF_X=1
run SNRF_LIB_PATH=$PWD/libsnrf_b4.so
run SNRF_X=1 --brick-gb 10
run SNRF_LIB_PATH=$PWD/libsnrf_b4.so --brick-gb 10
echo "== b4 variant parity"; SNRF_LIB_PATH=$PWD/libsnrf_b4.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_properties.py -m gpu -q -x > gpurun_out/c13_tests_b4.log 2>&1; tail -3 gpurun_out/c13_tests_b4.log
