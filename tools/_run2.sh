mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_backward.py tests/test_gpu_shapes.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/s12_pytest.log
cat gpurun_out/s12_pytest.log
(TGGCN_RECURRENT_MODE=2 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "dot" 2>&1 | tail -3) > gpurun_out/s12_pytest_big.log
cat gpurun_out/s12_pytest_big.log
