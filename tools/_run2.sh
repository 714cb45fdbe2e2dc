mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_fullsize.py tests/test_gpu_bf16.py tests/test_gpu_train_loop.py tests/test_gpu_linear.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/s6_pytest.log
cat gpurun_out/s6_pytest.log
timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s6_train.txt 2>&1; tail -6 gpurun_out/s6_train.txt
