#!/bin/bash
# Round-2 call 5: compile-time brick counts (straight-line level loops); occupancy / L1-allocation variants; ncu of the
# bucketed feature kernel and the output-layer GEMM.
mkdir -p gpurun_out
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c5_tests.log 2>&1; tail -4 gpurun_out/c5_tests.log
. tools/run_fn.sh
run SNRF_X=1 --feature-cutoff 5.96e-8
run SNRF_X=1 --feature-cutoff 5.96e-8 --brick-gb 10
run SNRF_LIB_PATH=$PWD/libsnrf_na.so --feature-cutoff 5.96e-8
run SNRF_LIB_PATH=$PWD/libsnrf_m4.so --feature-cutoff 5.96e-8
run SNRF_LIB_PATH=$PWD/libsnrf_m2.so --feature-cutoff 5.96e-8
run SNRF_X=1 --feature-cutoff 5.96e-8 --chunk 131072
for K in march_kernel sam_bucket_kernel tapgemm_kernel; do
echo "== ncu $K"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:$K -s 40 -c 1 -f -o gpurun_out/c5_$K \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline --feature-cutoff 5.96e-8 > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/c5_$K.ncu-rep > gpurun_out/c5_${K}_ncu.txt 2>&1; cat gpurun_out/c5_${K}_ncu.txt
python tools/ncu_opmix.py gpurun_out/c5_$K.ncu-rep 30 >> gpurun_out/c5_${K}_ncu.txt 2>&1
python tools/ncu_lines.py gpurun_out/c5_$K.ncu-rep 40 >> gpurun_out/c5_${K}_ncu.txt 2>&1
done
