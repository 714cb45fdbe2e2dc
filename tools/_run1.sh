mkdir -p gpurun_out
timeout 300 python -c 'import __graft_entry__ as g; g.smoke()' > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 1200 python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; tail -c 400 gpurun_out/r02_bench_final.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference.json 2> gpurun_out/r02_bench_reference.err
python - <<'PY'
import json
for f in ('gpurun_out/r02_bench_final.json','gpurun_out/r02_bench_reference.json'):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f, d.get('value'), d.get('ms_per_step'), d.get('e2e'), d.get('roofline'), d.get('gpu_launches'))
        for k in ('train_step','train_step_bf16'):
            if k in d: print(k, d[k])
    except Exception as e: print(f, 'ERR', e)
PY
