#!/bin/bash
# Round-2 call 23 (1 GPU): swizzled 64-byte a_tile (3 CTAs fit the 132 KB carve-out) - default carve-out vs forced 164 KB.
. tools/run_fn.sh
run SNRF_X=0
run SNRF_MARCH_CARVEOUT=71
run SNRF_MARCH_CARVEOUT=57
run SNRF_X=0
