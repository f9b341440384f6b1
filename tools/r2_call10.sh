#!/bin/bash
# Round-2 call 10 (2 GPUs): fp16 rows / march-first / push exchange correctness, then N = 2 bench variants.
mkdir -p gpurun_out
echo "== 1-GPU suite"; timeout 900 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py > gpurun_out/c10_tests.log 2>&1; tail -4 gpurun_out/c10_tests.log
echo "== multi-GPU tests"; timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/c10_tests_mgpu.log 2>&1; tail -6 gpurun_out/c10_tests_mgpu.log
runn() { n=$1; tag=$2; shift 2; echo "== N=$n $tag $*"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/c10_n${n}_$tag.json 2> gpurun_out/c10_n${n}_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c10_n${n}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "chunk", d["config"]["chunk"], d["config"]["exchange_check"] is not None, {k: round(v,3) for k,v in d["roofline"]["kernel_share_ms_per_step"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c10_n${n}_$tag.err").read()[-1500:])
PY
}
runn 2 dma_f16 --gather dma
runn 2 push_f16 --gather push
runn 2 push_f32 --gather push --feature-dtype f32
runn 2 dma_f16_mf --gather dma --march-first 1
echo "== 1 GPU default"; timeout 300 python bench.py --no-cpu-baseline > gpurun_out/c10_bench.json 2> gpurun_out/c10_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c10_bench.json')); print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_share_ms_per_step'])"
echo "== 1 GPU f16 rows"; timeout 300 python bench.py --no-cpu-baseline --feature-dtype f16 > gpurun_out/c10_bench_f16.json 2> gpurun_out/c10_bench_f16.err; python -c "
import json; d=json.load(open('gpurun_out/c10_bench_f16.json')); print(round(d['value'],2), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), d['roofline']['kernel_share_ms_per_step'])"
