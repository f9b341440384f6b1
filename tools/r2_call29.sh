#!/bin/bash
mkdir -p gpurun_out
timeout 25 python -m pytest tests/test_mask_decoder.py -m gpu -q > gpurun_out/c29_tests.log 2>&1; tail -12 gpurun_out/c29_tests.log
