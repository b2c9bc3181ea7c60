#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "expanded or insert or spmv_all" > gpurun_out/pytest_exp.log 2>&1; tail -25 gpurun_out/pytest_exp.log
