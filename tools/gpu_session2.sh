#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
for x in 1 0; do MGP_XFFT=$x python tools/xfft_probe.py 512 5; done 2>&1 | tee gpurun_out/s2_probe.txt
MGP_FORCE_SLAB=0 python tools/xfft_probe.py 512 5 2>&1 | tee -a gpurun_out/s2_probe.txt
# launch list with device times: fused path and the cuFFT 1-D + transpose path
for x in 1 0; do
  MGP_XFFT=$x ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s2_launches_xf$x.csv python tools/xfft_probe.py 512 1 > /dev/null 2>&1
done
# full capture of the fused kernels
ncu --set full --clock-control none --import-source on -k regex:k_xfft -c 2 -f -o gpurun_out/s2_xfft python tools/xfft_probe.py 512 1 > gpurun_out/s2_ncu.log 2>&1
tail -3 gpurun_out/s2_ncu.log
ls -la gpurun_out/s2_*
