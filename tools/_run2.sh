mkdir -p gpurun_out
for plan in 16 32; do
TGGCN_CL_PLAN=$plan timeout 300 python tools/profile_stages.py > gpurun_out/s10_stages_$plan.txt 2>&1; echo "plan $plan"; grep -E "forward|bigru" gpurun_out/s10_stages_$plan.txt
done
TGGCN_BIGRU_CLUSTER=0 timeout 300 python tools/profile_stages.py > gpurun_out/s10_stages_res.txt 2>&1; echo resident; grep -E "forward|bigru" gpurun_out/s10_stages_res.txt
(TGGCN_CL_PLAN=32 timeout 600 python -m pytest tests/test_gpu_bigru.py tests/test_gpu_parity.py -m gpu -q -x 2>&1 | tail -3) > gpurun_out/s10_pytest.log; cat gpurun_out/s10_pytest.log
