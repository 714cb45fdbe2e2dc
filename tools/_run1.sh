mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) > gpurun_out/s34_pytest_gpu.log; tail -2 gpurun_out/s34_pytest_gpu.log
timeout 300 python tools/profile_stages.py --iters 5 2>&1 | tail -13
