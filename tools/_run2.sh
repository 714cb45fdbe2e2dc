mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/s17_pytest.log
cat gpurun_out/s17_pytest.log
timeout 300 python tools/sweep_configs.py --what bimanual > gpurun_out/s17_sweep.txt 2>&1; grep bimanual gpurun_out/s17_sweep.txt
