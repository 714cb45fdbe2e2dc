mkdir -p gpurun_out
run() { echo "== $1"; TGGCN_NVCC_DEFS="$1" python 2g-gcn_b200/build.py --force > /dev/null 2>&1 || echo BUILD FAILED; timeout 100 python tools/profile_stages.py 2>&1 | grep -E "forward|bigru|segment"; }
{
run ""
run "-DSEG_NO_SHADOW"
run "-DGRID_FENCE_BARRIER"
run "-DSEG_EXP_ACCURATE"
} > gpurun_out/s8_ab.txt 2>&1
TGGCN_NVCC_DEFS="" python 2g-gcn_b200/build.py --force > /dev/null 2>&1
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> gpurun_out/s8_ab.txt
cat gpurun_out/s8_ab.txt
