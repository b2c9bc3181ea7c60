mkdir -p gpurun_out/r2_55
python -m pytest tests -m gpu -x -q > gpurun_out/r2_55/pytest.log 2>&1; tail -12 gpurun_out/r2_55/pytest.log
