#!/bin/bash
# Round-2 call 4: march kernel with oracle-mirroring arithmetic + bricks: whole GPU suite, diag, occupancy variants.
mkdir -p gpurun_out
echo "== diag"; timeout 600 python tools/diag_round2.py march_new_vs_v1 > gpurun_out/c4_diag.log 2>&1; grep "^\[" gpurun_out/c4_diag.log | sort -u; tail -2 gpurun_out/c4_diag.log
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_tests.log 2>&1; tail -12 gpurun_out/c4_tests.log
echo "== GPU tests, bucketed"; SNRF_FEATURE_CUTOFF=5.96e-8 timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c4_tests_bucketed.log 2>&1; tail -6 gpurun_out/c4_tests_bucketed.log
. tools/run_fn.sh
run SNRF_X=1 --feature-cutoff 5.96e-8
run SNRF_LIB_PATH=$PWD/libsnrf_m4.so --feature-cutoff 5.96e-8
run SNRF_X=1 --feature-cutoff 5.96e-8 --chunk 131072
run SNRF_LIB_PATH=$PWD/libsnrf_m4.so --feature-cutoff 5.96e-8 --chunk 131072
run SNRF_LIB_PATH=$PWD/libsnrf_m4.so --feature-cutoff 5.96e-8 --chunk 131072 --brick-gb 10
