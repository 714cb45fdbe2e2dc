#!/bin/bash
# One pass over everything profiles/ keeps for a round (run on the GPU box through gpurun):  bash tools/evidence_run.sh <tag>
tag=${1:-rXX}
o=gpurun_out
mkdir -p $o
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5) > $o/${tag}_pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > $o/${tag}_bench_n1.json 2> $o/${tag}_bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/${tag}_bench_reference.json 2> $o/${tag}_bench_reference.err
timeout 300 python tools/profile_stages.py > $o/${tag}_stages.txt 2>&1
timeout 300 python tools/profile_train.py --iters 5 > $o/${tag}_train_phases.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train > $o/${tag}_launches_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/${tag}_launches_train.csv python tools/profile_train.py --iters 1 > $o/${tag}_launches_train.log 2>&1
timeout 600 python -c 'import __graft_entry__ as g; g.smoke()' > $o/${tag}_smoke.log 2>&1
for k in segment_kernel bigru_res_kernel geo_gcn_kernel frame_messages_kernel heads_kernel; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 1 -c 1 -o $o/${tag}_$k python tools/profile_stages.py --iters 1 > $o/${tag}_ncu_$k.log 2>&1
done
# the hoisted segment W_ih projection group: the 11th gemm_tc launch of the two forwards profile_stages runs
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -s 11 -c 1 -o $o/${tag}_gemm_tc python tools/profile_stages.py --iters 1 > $o/${tag}_ncu_gemm_tc.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:geo_gcn_bwd_kernel -s 1 -c 1 -o $o/${tag}_geo_gcn_bwd_kernel python tools/profile_train.py --iters 1 > $o/${tag}_ncu_geo_gcn_bwd.log 2>&1
tail -2 $o/${tag}_pytest_gpu.log; tail -3 $o/${tag}_smoke.log
cat $o/${tag}_stages.txt
cat $o/${tag}_train_phases.txt | tail -12
