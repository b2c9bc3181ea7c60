mkdir -p gpurun_out/r2_49
python -m pytest tests -m gpu -x -q > gpurun_out/r2_49/pytest.log 2>&1; tail -3 gpurun_out/r2_49/pytest.log
