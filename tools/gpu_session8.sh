#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
export PROBE_ONLY=1,6,8,9
for v in "1 1 2" "1 1 1" "1 0 2" "1 0 1" "0 0 2"; do set -- $v
  echo "== XFFT=$1 DMA=$2 CPS=$3"; MGP_XFFT=$1 MGP_XFFT_DMA=$2 MGP_XFFT_CPS=$3 timeout 200 $TR tools/exchange_probe.py 512 8 2>&1 | grep -E "N=|rror"
done | tee gpurun_out/s8_exchange_2gpu.txt
B="bench.py --gpus 2 --nmesh 512 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline"
for v in "1 1 2" "1 0 1" "1 1 1"; do set -- $v
  MGP_XFFT=$1 MGP_XFFT_DMA=$2 MGP_XFFT_CPS=$3 timeout 300 $TR $B > gpurun_out/s8_bench2_xf$1_dma$2_cps$3.json 2> gpurun_out/s8_bench2_xf$1_dma$2_cps$3.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s8_bench2_*.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm")})
    except Exception as e: print(f, "failed", e)
PY
