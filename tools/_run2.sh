mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_linear.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/s13_pytest.log; cat gpurun_out/s13_pytest.log
timeout 300 python tools/profile_stages.py > gpurun_out/s13_stages.txt 2>&1; grep -E "forward|gemm" gpurun_out/s13_stages.txt
timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s13_train.txt 2>&1; grep -E "forward|backward|train step" gpurun_out/s13_train.txt
