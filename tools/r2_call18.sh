#!/bin/bash
# Round-2 call 18 (2 GPUs): the TMA-driven push kernel - multi-GPU bit-identity tests under both push kernels, 2-rank bench.
mkdir -p gpurun_out
T=c18
echo "== multi-GPU tests (push = tma)"; timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -k "push" > gpurun_out/${T}_tests_mgpu.log 2>&1; tail -3 gpurun_out/${T}_tests_mgpu.log
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
runn 2 tma_nomf --gather push --march-first 0
SNRF_PUSH=st runn 2 st_nomf --gather push --march-first 0
