mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/s2_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/s2_bench_n1.json 2> gpurun_out/s2_bench_n1.err
timeout 300 python tools/profile_stages.py > gpurun_out/s2_stages.txt 2>&1
timeout 300 python tools/profile_train.py --iters 5 > gpurun_out/s2_train_phases.txt 2>&1
tail -3 gpurun_out/s2_pytest_gpu.log; cat gpurun_out/s2_stages.txt | tail -30; tail -25 gpurun_out/s2_train_phases.txt
