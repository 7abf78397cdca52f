#!/bin/bash
# First thing to run next round on 8 GPUs (gpurun --gpus 8): what the fused slab path does at the target scale.
#   1. exchange pieces alone at 512^3 and 1024^3 (per-direction NVLink rate with 7/8 of the slab remote)
#   2. whole-step benches at 512^3 (the weak-scaling point) and 1024^3 (the north-star target) for the exchange engines
#   3. the 8-rank parity tests
# Every step is bounded; the whole script stays under ~6 minutes of box time.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533"
for N in 512 1024; do
  PROBE_ONLY=0,1,2,3,4,8,9 timeout 120 $TR tools/exchange_probe.py $N 8 2>&1 | grep -E "N=|rror"
done | tee gpurun_out/n8_exchange.txt
for N in 512 1024; do
  MGP_XFFT_WIDE=1 PROBE_ONLY=1,2,8 timeout 120 $TR tools/exchange_probe.py $N 8 2>&1 | grep -E "N=|rror"
done | tee gpurun_out/n8_exchange_wide.txt
B="bench.py --gpus 8 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
MGP_XFFT_WIDE=1 timeout 120 $TR $B --nmesh 512 > gpurun_out/n8_bench_512_wide.json 2> gpurun_out/n8_bench_512_wide.err
for v in "1 0 1" "1 0 2" "0 0 1" "1 1 1"; do set -- $v
  MGP_XFFT=$1 MGP_XFFT_DMA=$2 MGP_XFFT_CPS=$3 timeout 120 $TR $B --nmesh 512 > gpurun_out/n8_bench_512_xf$1_dma$2_cps$3.json 2> gpurun_out/n8_bench_512_xf$1_dma$2_cps$3.err
done
for v in "1 0 1" "0 0 1"; do set -- $v
  MGP_XFFT=$1 MGP_XFFT_DMA=$2 MGP_XFFT_CPS=$3 timeout 200 $TR $B --nmesh 1024 > gpurun_out/n8_bench_1024_xf$1_dma$2_cps$3.json 2> gpurun_out/n8_bench_1024_xf$1_dma$2_cps$3.err
done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/n8_bench_*.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, "ms/step %.3f  value %.3e  step frac %.2f" % (d["ms_per_step"], d["value"], d["roofline"]["step"]["frac"]),
              {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm","MoveParticles","Sort")})
    except Exception as e: print(f, "failed", e)
PY
timeout 600 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "8-fofr or 8-lcdm or 8-dgp or (matches_reference and 8) or (on_slabs and 8)" > gpurun_out/n8_mgpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/n8_mgpu_tests.log; tail -5 gpurun_out/n8_mgpu_tests.log
