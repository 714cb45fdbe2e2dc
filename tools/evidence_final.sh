#!/bin/bash
# Final round-2 evidence pass (gpurun):  bash tools/evidence_final.sh
o=gpurun_out
mkdir -p $o
timeout 1200 python bench.py --steps 20 --warmup 5 > $o/r02_bench_final.json 2> $o/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $o/r02_bench_reference.json 2> $o/r02_bench_reference.err
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > $o/r02_smoke.log 2>&1
timeout 900 python tools/sweep_configs.py > $o/r02_sweep_final.txt 2>&1
timeout 300 python tools/profile_stages.py --iters 5 > $o/r02_stages_final.txt 2>&1
timeout 300 python tools/profile_train.py --iters 5 > $o/r02_train_phases.txt 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/r02_launches_final.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-train --no-extras > $o/r02_launches_bench.log 2>&1
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $o/r02_launches_train_final.csv python tools/profile_train.py --iters 1 > $o/r02_launches_train.log 2>&1
tail -3 $o/r02_smoke.log; tail -c 300 $o/r02_bench_final.err; cat $o/r02_sweep_final.txt | tail -25
