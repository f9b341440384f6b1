#!/bin/bash
# Round-2 call 6: the rewritten bench.py (three configs, per-kernel roofline, library-side D2H for e2e) + GPU suite.
mkdir -p gpurun_out
echo "== GPU tests"; timeout 900 python -m pytest tests -m gpu -q > gpurun_out/c6_tests.log 2>&1; tail -4 gpurun_out/c6_tests.log
echo "== bench (default)"; timeout 600 python bench.py > gpurun_out/c6_bench.json 2> gpurun_out/c6_bench.err; tail -c 1500 gpurun_out/c6_bench.json; tail -3 gpurun_out/c6_bench.err
echo "== bench rgb"; timeout 300 python bench.py --config rgb --no-cpu-baseline > gpurun_out/c6_bench_rgb.json 2>> gpurun_out/c6_bench.err; tail -c 700 gpurun_out/c6_bench_rgb.json
echo "== bench clipseg_patch"; timeout 300 python bench.py --config clipseg_patch --no-cpu-baseline > gpurun_out/c6_bench_clipseg.json 2>> gpurun_out/c6_bench.err; tail -c 900 gpurun_out/c6_bench_clipseg.json
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/c6_bench_ref.json 2>> gpurun_out/c6_bench.err; tail -c 500 gpurun_out/c6_bench_ref.json
tail -5 gpurun_out/c6_bench.err
