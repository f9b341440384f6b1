#!/bin/bash
mkdir -p gpurun_out
timeout 80 python bench.py > gpurun_out/f2_bench.json 2> gpurun_out/f2_bench.err; tail -c 600 gpurun_out/f2_bench.json
