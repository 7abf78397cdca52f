#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for v in "1 1" "1 0" "0 0"; do set -- $v
  echo "== XFFT=$1 DMA=$2"; MGP_XFFT=$1 MGP_XFFT_DMA=$2 python tools/exchange_probe.py 512 8 2>&1 | grep -E "N=|rror|unavail"
done | tee gpurun_out/s7_exchange_1gpu.txt
echo "== XFFT=1 DMA=1 TRIM=0"; MGP_XFFT_TRIM=0 python tools/exchange_probe.py 512 8 2>&1 | grep -E "pipeline" | tee -a gpurun_out/s7_exchange_1gpu.txt
echo "== XFFT=1 DMA=1 TRIM=16"; MGP_XFFT_TRIM=16 python tools/exchange_probe.py 512 8 2>&1 | grep -E "pipeline|fused" | tee -a gpurun_out/s7_exchange_1gpu.txt
