mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullsize.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/s19_pytest.log
cat gpurun_out/s19_pytest.log
timeout 600 python tools/sweep_configs.py --what cad120 --batches 32 2>&1 | grep cad120
timeout 600 python tools/sweep_configs.py --what mphoi 2>&1 | grep "B=32"
