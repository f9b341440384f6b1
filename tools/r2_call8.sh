#!/bin/bash
# Round-2 call 8: TMA-fed output-layer GEMM: correctness against the round-1 kernel, suite, A/B bench, ncu.
mkdir -p gpurun_out
echo "== tapgemm check"; timeout 600 python tools/check_tapgemm.py 2>&1 | tail -20
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/c8_tests.log 2>&1; tail -4 gpurun_out/c8_tests.log
. tools/run_fn.sh
run SNRF_TAPGEMM=v1
run SNRF_TAPGEMM=tma
echo "== ncu tapgemm_tma"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tapgemm_tma_kernel -s 8 -c 1 -f -o gpurun_out/c8_tapgemm_tma \
  python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python tools/ncu_summary.py gpurun_out/c8_tapgemm_tma.ncu-rep > gpurun_out/c8_tapgemm_tma_ncu.txt 2>&1; cat gpurun_out/c8_tapgemm_tma_ncu.txt
