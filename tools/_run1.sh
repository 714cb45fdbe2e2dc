mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/s11_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/s11_bench_n1.json 2> gpurun_out/s11_bench_n1.err
timeout 900 python tools/sweep_configs.py > gpurun_out/s11_sweep.txt 2>&1
tail -3 gpurun_out/s11_pytest_gpu.log; cat gpurun_out/s11_sweep.txt; tail -c 600 gpurun_out/s11_bench_n1.err
