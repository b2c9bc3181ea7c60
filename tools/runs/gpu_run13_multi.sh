#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s > gpurun_out/pytest_multi.log 2>&1; tail -5 gpurun_out/pytest_multi.log
