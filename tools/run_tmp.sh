mkdir -p gpurun_out/r2_37
python -m pytest tests -m gpu -x -q -k "partition or strategies or poisson_p1" > gpurun_out/r2_37/pytest.log 2>&1; tail -12 gpurun_out/r2_37/pytest.log
