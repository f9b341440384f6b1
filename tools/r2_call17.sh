#!/bin/bash
mkdir -p gpurun_out
T=c17
runn() { n=$1; tag=$2; shift 2; echo "== N=$n $tag $*"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${T}_n${n}_$tag.json 2> gpurun_out/${T}_n${n}_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_n${n}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "chunk", d["config"]["chunk"], d["config"]["exchange_check"] is not None, {k: round(v,3) for k,v in d["roofline"]["kernel_share_ms_per_step"].items()})
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${T}_n${n}_$tag.err").read()[-1500:])
PY
}
runn 8 push_nomf --gather push --march-first 0
runn 8 push_mf --gather push --march-first 1
SNRF_PUSH_GRID=128 runn 8 push_nomf_g128 --gather push --march-first 0
runn 8 push_nomf_10k --gather push --march-first 0 --chunk 10240
runn 8 dma_nomf --gather dma --march-first 0
