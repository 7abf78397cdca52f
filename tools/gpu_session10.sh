#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
PROBE_ONLY=1,2,8 timeout 60 $TR tools/exchange_probe.py 512 8 2>&1 | grep -E "N=|rror" | tee gpurun_out/s10_exchange_2gpu.txt
timeout 70 $TR bench.py --gpus 2 --nmesh 512 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/s10_bench2_512.json 2> gpurun_out/s10_bench2_512.err
python - <<'PY'
import json
try:
    d=json.loads([x for x in open("gpurun_out/s10_bench2_512.json") if x.startswith('{')][-1])
    print("2 GPUs 512: ms/step %.3f" % d["ms_per_step"], {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm")})
except Exception as e: print("failed", e)
PY
