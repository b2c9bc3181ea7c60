#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "interior_facets or golden or exterior" > gpurun_out/pytest_dS.log 2>&1; tail -25 gpurun_out/pytest_dS.log
