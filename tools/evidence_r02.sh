#!/bin/bash
# Round-2 evidence pass (run on the GPU box through gpurun):  bash tools/evidence_r02.sh
# launch lists of the bench command and of a training step, ncu --set full captures of the kernels DESIGN.md §4 / §6 quote.
o=gpurun_out
mkdir -p $o
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $o/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $o/r02_launches_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/r02_launches_train.csv python tools/profile_train.py --iters 1 > $o/r02_launches_train.log 2>&1
cap() {  # name, kernel regex, skip, command...
  local name=$1 k=$2 skip=$3; shift 3
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s $skip -c 1 -f -o $o/r02_$name "$@" > $o/r02_ncu_$name.log 2>&1
}
cap segment_kernel segment_kernel 1 python tools/profile_stages.py --iters 1
cap bigru_res_kernel bigru_res_kernel 1 python tools/profile_stages.py --iters 1
# projection kernel: 6 gemm16 launches per forward; the 6th is the hoisted segment W_ih stage (gemm_gs); second forward
cap gemm16_gs gemm16_kernel 11 python tools/profile_stages.py --iters 1
cap gemm16_embed gemm16_kernel 6 python tools/profile_stages.py --iters 1
cap pack16x pack16x_kernel 11 python tools/profile_stages.py --iters 1
cap frame_messages_kernel frame_messages_kernel 1 python tools/profile_stages.py --iters 1
# large-batch path: CAD-120 B=256, 8 steps: cell GEMM (GRU mode with an input part) = every 3rd step_tc launch after the BiGRU's
cap step_cell step_tc_kernel 40 python tools/run_once.py --shape cad120 --B 256 --T 8 --D 512 --iters 1
cap seg_attend seg_attend_kernel 4 python tools/run_once.py --shape cad120 --B 256 --T 8 --D 512 --iters 1
cap bigru_cluster bigru_cluster_kernel 0 python tools/run_once.py --shape cad120 --B 8 --T 128 --D 512 --iters 1
cap segment_bwd segment_bwd_kernel 0 python tools/profile_train.py --iters 1
ls -la $o/r02_*.ncu-rep | awk '{print $5, $9}'
