#!/bin/bash
# last seconds of the round-1 GPU budget: parity of the round-2 experiment variants (chunk_walk phase 2 + padded lists)
mkdir -p gpurun_out
BFX_CHUNK_DIET=1 BFX_CHUNKS_PAD4=1 timeout 25 python -m pytest tests/test_gpu_parity.py -x -q -k "test_poisson_p1_matrix or test_poisson_p1_32 or test_poisson_p2" > gpurun_out/r55_pytest.log 2>&1; tail -3 gpurun_out/r55_pytest.log
