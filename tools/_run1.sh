mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "backward or train or dropin or bf16 or parity" 2>&1 | tail -3)
timeout 300 python tools/host_overhead.py --train 2>&1 | grep -E "host enqueue|device time"
timeout 300 python tools/host_overhead.py 2>&1 | grep -E "host enqueue|device time"
