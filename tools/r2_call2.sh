#!/bin/bash
# Round-2 call 2: new march kernel (lane = sample) correctness + A/B against the round-1 kernel, diagnostics of the
# backward failures, one ncu capture of the new march kernel.
mkdir -p gpurun_out
echo "== GPU tests (new march kernel)"; timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c2_tests.log 2>&1; tail -15 gpurun_out/c2_tests.log
echo "== unverified GPU tests"; SNRF_RUN_UNVERIFIED=1 timeout 900 python -m pytest tests -m "gpu and hw_unverified" -q > gpurun_out/c2_tests_unverified.log 2>&1; tail -15 gpurun_out/c2_tests_unverified.log
echo "== bucketed suite"; SNRF_FEATURE_CUTOFF=5.96e-8 timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c2_tests_bucketed.log 2>&1; tail -5 gpurun_out/c2_tests_bucketed.log
echo "== diag"; timeout 600 python tools/diag_round2.py > gpurun_out/c2_diag.log 2>&1; grep "^\[" gpurun_out/c2_diag.log
. tools/run_fn.sh
run SNRF_MARCH=v1 --feature-cutoff 5.96e-8
run SNRF_MARCH=new --feature-cutoff 5.96e-8
run SNRF_MARCH=new
echo "== ncu march"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 40 -c 1 -f -o gpurun_out/c2_march \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --feature-cutoff 5.96e-8 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/c2_march.ncu-rep > gpurun_out/c2_march_ncu.txt 2>&1; cat gpurun_out/c2_march_ncu.txt
python tools/ncu_opmix.py gpurun_out/c2_march.ncu-rep 30 >> gpurun_out/c2_march_ncu.txt 2>&1
python tools/ncu_lines.py gpurun_out/c2_march.ncu-rep 40 >> gpurun_out/c2_march_ncu.txt 2>&1
