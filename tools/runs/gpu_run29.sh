#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "matrix_free" > gpurun_out/pytest_mf.log 2>&1; tail -30 gpurun_out/pytest_mf.log
