mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_backward.py tests/test_gpu_fullsize.py tests/test_gpu_bigru.py tests/test_gpu_train_loop.py tests/test_gpu_bf16.py -m gpu -q -x 2>&1 | tail -6) > gpurun_out/s14_pytest.log
cat gpurun_out/s14_pytest.log
timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s14_train.txt 2>&1; grep -E "forward|backward|train step" gpurun_out/s14_train.txt
TGGCN_BWD_RES=0 timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s14_train_old.txt 2>&1; grep -E "forward|backward|train step" gpurun_out/s14_train_old.txt
