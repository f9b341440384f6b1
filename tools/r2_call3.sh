#!/bin/bash
# Round-2 call 3: brick layout (cell-major grid levels) for the march kernel - correctness, brick-budget sweep, chunk
# size effect, diagnostics of the per-sample density differences of the lane = sample kernel.
mkdir -p gpurun_out
echo "== diag"; timeout 600 python tools/diag_round2.py march_new_vs_v1 > gpurun_out/c3_diag.log 2>&1; grep "^\[" gpurun_out/c3_diag.log; tail -3 gpurun_out/c3_diag.log
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c3_tests.log 2>&1; tail -12 gpurun_out/c3_tests.log
. tools/run_fn.sh
for gb in 0 0.02 1.5 4 10 25 65; do run SNRF_X=1 --feature-cutoff 5.96e-8 --brick-gb $gb; done
run SNRF_X=1 --feature-cutoff 5.96e-8 --brick-gb 10 --chunk 131072
run SNRF_X=1 --feature-cutoff 5.96e-8 --brick-gb 10 --chunk 640000
echo "== ncu march (10 GB bricks)"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 40 -c 1 -f -o gpurun_out/c3_march \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --feature-cutoff 5.96e-8 --brick-gb 10 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/c3_march.ncu-rep > gpurun_out/c3_march_ncu.txt 2>&1; cat gpurun_out/c3_march_ncu.txt
python tools/ncu_opmix.py gpurun_out/c3_march.ncu-rep 30 >> gpurun_out/c3_march_ncu.txt 2>&1
python tools/ncu_lines.py gpurun_out/c3_march.ncu-rep 40 >> gpurun_out/c3_march_ncu.txt 2>&1
