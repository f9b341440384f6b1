#!/bin/bash
mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_march_variants.py -m gpu -q > gpurun_out/c25_tests.log 2>&1; tail -15 gpurun_out/c25_tests.log
