mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/s30_pytest_gpu.log
timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s30_train.txt 2>&1
timeout 300 python tools/profile_stages.py --iters 5 > gpurun_out/s30_stages.txt 2>&1
tail -3 gpurun_out/s30_pytest_gpu.log; grep -v Warn gpurun_out/s30_train.txt | tail -8; tail -20 gpurun_out/s30_stages.txt
