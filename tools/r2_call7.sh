#!/bin/bash
# Round-2 call 7 (2 GPUs): two-rank exchange tests, bench at N = 2 (copy engines / multicast), plus 1-GPU rgb config.
mkdir -p gpurun_out
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/c7_tests_mgpu.log 2>&1; tail -5 gpurun_out/c7_tests_mgpu.log
for g in dma mc; do
echo "== bench N=2 $g"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --gather $g > gpurun_out/c7_bench_n2_$g.json 2> gpurun_out/c7_bench_n2_$g.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c7_bench_n2_$g.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), d["config"]["tiles"], d["config"]["exchange_check"], d["roofline"]["kernel_share_ms_per_step"])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c7_bench_n2_$g.err").read()[-1500:])
PY
done
echo "== bench rgb (1 GPU)"; timeout 300 python bench.py --config rgb --no-cpu-baseline > gpurun_out/c7_bench_rgb.json 2> gpurun_out/c7_bench_rgb.err; tail -c 600 gpurun_out/c7_bench_rgb.json; tail -3 gpurun_out/c7_bench_rgb.err
echo "== new golden tests (1 GPU)"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/c7_tests_parity.log 2>&1; tail -5 gpurun_out/c7_tests_parity.log
