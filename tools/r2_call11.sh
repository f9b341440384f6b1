#!/bin/bash
# Round-2 call 11 (8 GPUs): exchange variants with fp16 rows.
mkdir -p gpurun_out
echo "== 1-GPU suite (W1 bulk staging)"; CUDA_VISIBLE_DEVICES=0 timeout 600 python -m pytest tests -m gpu -q -x --deselect tests/test_multi_gpu.py > gpurun_out/c11_tests.log 2>&1; tail -3 gpurun_out/c11_tests.log
runn() { n=$1; tag=$2; shift 2; echo "== N=$n $tag $*"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/c11_n${n}_$tag.json 2> gpurun_out/c11_n${n}_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c11_n${n}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "chunk", d["config"]["chunk"], d["config"]["exchange_check"] is not None, {k: round(v,3) for k,v in d["roofline"]["kernel_share_ms_per_step"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c11_n${n}_$tag.err").read()[-1500:])
PY
}
runn 8 dma_mf --gather dma
runn 8 dma_nomf --gather dma --march-first 0
runn 8 push_mf --gather push
runn 8 dma_mf_27k --gather dma --chunk 27008
SNRF_DMA_SPLIT=2 runn 8 dma_mf_split2 --gather dma
runn 4 dma_mf --gather dma
