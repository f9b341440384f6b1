#!/bin/bash
# First gpurun call of the next round: everything that was written after the round-1 GPU budget was spent.
#   gpurun --timeout 1500 -- 'bash tools/round2_first_call.sh'
# Writes gpurun_out/r2_*.log / *.json.  Each step has its own timeout so that one hang cannot eat the call.
mkdir -p gpurun_out
echo "== verified GPU tests"; timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_tests_verified.log 2>&1; tail -3 gpurun_out/r2_tests_verified.log
echo "== hardware-unverified GPU tests (no -x: see every failure)"
SNRF_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests -m "gpu and hw_unverified" -q > gpurun_out/r2_tests_unverified.log 2>&1; tail -25 gpurun_out/r2_tests_unverified.log
echo "== the verified parity suite through the bucketed feature kernel (SNRF_FEATURE_CUTOFF=5.96e-8)"
SNRF_FEATURE_CUTOFF=5.96e-8 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_bucketed.log 2>&1; tail -8 gpurun_out/r2_tests_bucketed.log
echo "== ... and through early termination at 2^-24 (per-sample debug outputs of skipped samples read 0: stage tests may differ there)"
SNRF_EARLY_TERMINATION=5.96e-8 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/r2_tests_et.log 2>&1; tail -8 gpurun_out/r2_tests_et.log
echo "== bench (exact)"; timeout 400 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 600 gpurun_out/r2_bench.json
echo "== bench (early termination 1e-4)"; timeout 300 python bench.py --early-termination 1e-4 --no-cpu-baseline > gpurun_out/r2_bench_et.json 2>> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench_et.json
echo "== bench (bucketed feature kernel, cut-off 2^-24)"; timeout 300 python bench.py --feature-cutoff 5.96e-8 --no-cpu-baseline > gpurun_out/r2_bench_bucket.json 2>> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench_bucket.json
echo "== bench (bucketed feature kernel, exact zeros only)"; timeout 300 python bench.py --feature-cutoff 0 --no-cpu-baseline > gpurun_out/r2_bench_bucket0.json 2>> gpurun_out/r2_bench.err; tail -c 400 gpurun_out/r2_bench_bucket0.json
echo "== camera in front"; timeout 200 python tools/bench_camera.py > gpurun_out/r2_camera.json 2> gpurun_out/r2_camera.err; cat gpurun_out/r2_camera.json; tail -3 gpurun_out/r2_camera.err
echo "== training step"; timeout 300 python tools/bench_train.py > gpurun_out/r2_train.json 2> gpurun_out/r2_train.err; cat gpurun_out/r2_train.json; tail -3 gpurun_out/r2_train.err
