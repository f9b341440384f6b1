#!/bin/bash
# Round-2 call 24 (8 GPUs): multicast push against the peer-pointer push, march-first and interleaved.
mkdir -p gpurun_out
T=c24
runn() { n=$1; tag=$2; shift 2; echo "== N=$n $tag $*"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/${T}_n${n}_$tag.json 2> gpurun_out/${T}_n${n}_$tag.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${T}_n${n}_$tag.json").read().strip().splitlines()[-1])
    print(round(d["value"],2), round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "chunk", d["config"]["chunk"], d["config"]["exchange_check"] is not None, {k: round(v,3) for k,v in d["roofline"]["kernel_share_ms_per_step"].items()})
    print(d["config"]["tiles"][:200])
except Exception as e:
    print("FAILED", e); print(open("gpurun_out/${T}_n${n}_$tag.err").read()[-1500:])
PY
}
runn 8 pushmc_mf --gather pushmc --march-first 1
runn 8 pushmc_nomf --gather pushmc --march-first 0
runn 8 push_mf --gather push --march-first 1
