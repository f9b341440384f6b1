#!/bin/bash
# Round-2 call 9 (8 GPUs): scaling of the tile exchange.
mkdir -p gpurun_out
runn() { n=$1; tag=$2; shift 2; echo "== N=$n $tag $*"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 "$@" > gpurun_out/c9_n${n}_$tag.json 2> gpurun_out/c9_n${n}_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/c9_n${n}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "chunk", d["config"]["chunk"], d["config"]["exchange_check"] is not None, {k: round(v,3) for k,v in d["roofline"]["kernel_share_ms_per_step"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/c9_n${n}_$tag.err").read()[-1200:])
PY
}
runn 8 dma --gather dma
runn 8 dma8k --gather dma --chunk 8192
runn 8 mc --gather mc
runn 4 dma --gather dma
runn 8 nccl --gather nccl
