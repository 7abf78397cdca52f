#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
timeout 300 $TR tools/exchange_probe.py 512 8 2>&1 | grep -E "N=|rror" | tee gpurun_out/s6_exchange_2gpu.txt
timeout 600 python -m pytest tests/test_slab_fused.py -x -q -m gpu > gpurun_out/s6_slab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s6_slab_tests.log; tail -4 gpurun_out/s6_slab_tests.log
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/s6_mgpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s6_mgpu_tests.log; tail -4 gpurun_out/s6_mgpu_tests.log
for N in 512 256; do for v in "1 1" "1 0" "0 0"; do set -- $v
  MGP_XFFT=$1 MGP_XFFT_DMA=$2 timeout 400 $TR bench.py --gpus 2 --nmesh $N --steps 5 --warmup 3 > gpurun_out/s6_bench2_xf$1_dma$2_$N.json 2> gpurun_out/s6_bench2_xf$1_dma$2_$N.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s6_bench2_*.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm")})
    except Exception as e: print(f, "failed", e)
PY
