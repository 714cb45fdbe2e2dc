mkdir -p gpurun_out
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "noise" 2>&1 | tail -3)
timeout 900 python bench.py --steps 20 --warmup 5 --no-train --no-extras --no-cpu-baseline > gpurun_out/s32_bench.json 2> gpurun_out/s32_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/s32_bench.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'])
PY
