mkdir -p gpurun_out/r2_32
python -m pytest tests -m gpu -x -q -k "q1 or Q1" > gpurun_out/r2_32/pytest_q1.log 2>&1
tail -15 gpurun_out/r2_32/pytest_q1.log
