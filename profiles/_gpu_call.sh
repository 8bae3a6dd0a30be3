mkdir -p gpurun_out/r2ang
(timeout 600 python -m pytest tests/test_angle_estimates_gpu.py tests/test_pipeline_gpu.py -m gpu -x -q 2>&1 | tail -15) > gpurun_out/r2ang/pytest.log
tail -6 gpurun_out/r2ang/pytest.log
