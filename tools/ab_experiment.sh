#!/bin/bash
# A/B builds on the GPU box (nvcc is in the image): rebuild the library with compile-time switches and time the recurrent stages.
#   bash tools/ab_experiment.sh            -> gpurun_out/ab.txt
# Switches: -DSEG_NO_SHADOW      segment kernel without the next phase's set-up between grid_arrive and grid_wait
#           -DGRID_FENCE_BARRIER grid barrier with two __threadfence instead of red.release / ld.acquire
# (profiles/r01_ab_barrier_shadow.txt also holds -DSEG_EXP_ACCURATE, a switch of the since-removed fused softmax.)
mkdir -p gpurun_out
run() { echo "== ${1:-default}"; TGGCN_NVCC_DEFS="$1" python 2g-gcn_b200/build.py --force > /dev/null 2>&1 || echo BUILD FAILED; timeout 100 python tools/profile_stages.py 2>&1 | grep -E "forward|bigru|segment"; }
{
run ""
run "-DSEG_NO_SHADOW"
run "-DGRID_FENCE_BARRIER"
} > gpurun_out/ab.txt 2>&1
TGGCN_NVCC_DEFS="" python 2g-gcn_b200/build.py --force > /dev/null 2>&1
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -4) >> gpurun_out/ab.txt
cat gpurun_out/ab.txt
