mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_linear.py -m gpu -q -x -k "linear16" 2>&1 | tail -15) > gpurun_out/s3_pytest_lin.log
cat gpurun_out/s3_pytest_lin.log
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_fullsize.py tests/test_gpu_bf16.py -m gpu -q -x 2>&1 | tail -15) > gpurun_out/s3_pytest.log
cat gpurun_out/s3_pytest.log
timeout 300 python tools/profile_stages.py > gpurun_out/s3_stages.txt 2>&1
TGGCN_GEMM16=0 timeout 300 python tools/profile_stages.py > gpurun_out/s3_stages_old.txt 2>&1
grep -E "forward|gemm" gpurun_out/s3_stages.txt gpurun_out/s3_stages_old.txt
