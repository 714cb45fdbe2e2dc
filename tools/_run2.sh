mkdir -p gpurun_out
(timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_backward.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/s18_pytest.log
cat gpurun_out/s18_pytest.log
timeout 300 python tools/profile_stages.py > gpurun_out/s18_stages.txt 2>&1; grep -E "forward|frame_msg|heads|geo_gcn" gpurun_out/s18_stages.txt
