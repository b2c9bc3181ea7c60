mkdir -p gpurun_out/r2_31
python -m pytest tests -m gpu -x -q -k "facet_functionals or two_fused or empty_inputs or interior_facets or exterior_facets or packed" > gpurun_out/r2_31/pytest_new.log 2>&1
tail -40 gpurun_out/r2_31/pytest_new.log
