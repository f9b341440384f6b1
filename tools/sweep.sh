#!/bin/bash
# usage: tools/sweep.sh lib1.so lib2.so ...   -> per-kernel ms/frame for each library variant
for lib in "$@"; do
  echo -n "$lib: "
  SNRF_LIB_PATH=$PWD/$lib timeout 200 python bench.py --steps 4 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
try:
    d=json.loads(sys.stdin.read()); k=d['roofline']['kernel_share_ms_per_step']
    print(f\"{d['value']:.2f} Mrays/s  ms/step {d['ms_per_step']:.2f}  march {k['march']:.2f} feature {k['feature']:.2f} gemm {k['tapgemm']:.2f}  e2e {d['e2e']['value']:.2f}\")
except Exception as e: print('FAILED', e)
"
done
