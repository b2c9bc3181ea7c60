#!/bin/bash
mkdir -p gpurun_out
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "chunked and p1_lex and True" > gpurun_out/racecheck_chunked.log 2>&1; tail -4 gpurun_out/racecheck_chunked.log
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "rowgather and 6-True" > gpurun_out/racecheck_rowgather.log 2>&1; tail -4 gpurun_out/racecheck_rowgather.log
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "spmv_assembled" > gpurun_out/racecheck_spmv.log 2>&1; tail -4 gpurun_out/racecheck_spmv.log
timeout 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "interior_facets or expanded or matrix_free" > gpurun_out/sanitizer_new.log 2>&1; tail -4 gpurun_out/sanitizer_new.log
