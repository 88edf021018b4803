#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_variants.py -m gpu -x -q -s 2>&1 | tail -12
