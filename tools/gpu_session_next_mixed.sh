#!/bin/bash
# Next round, 2 GPUs (gpurun --gpus 2): first GPU run of the mixed-radix fused x-transform (MGP_XFFT_MIXED=1, Nmesh = 320).
#   1. one rank, forced slab: r2c / c2r against the cuFFT 1-D + transpose path on the same context size (parity + time)
#   2. two ranks: exchange pieces, whole-step bench with and without it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
MGP_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_slab_fused.py -x -q -m gpu -k experimental 2>&1 | tail -5 | tee gpurun_out/mx_experimental_tests.txt
python - <<'PY' 2>&1 | tee gpurun_out/mx_parity.txt
import os, sys, numpy as np
sys.path.insert(0, ".")
os.environ["MGP_FORCE_SLAB"] = "1"
import mgpicola_b200 as mgp
for N in (320, 400):
    out = {}
    rng = np.random.default_rng(N)
    x = rng.standard_normal((N, N, N))
    for mixed in ("0", "1"):
        os.environ["MGP_XFFT_MIXED"] = mixed
        pm = mgp.PM(N, N, 100.0, grid_bytes=8)
        g = np.zeros((N + 1, N, 2 * (N // 2 + 1))); g[:N, :, :N] = x
        pm.upload_grid(mgp.GRID_DENSITY, g)
        pm.fft_r2c(mgp.GRID_DENSITY)
        k = pm.download_grid_k(mgp.GRID_DENSITY).copy()
        pm.fft_c2r(mgp.GRID_DENSITY)
        r = pm.download_grid(mgp.GRID_DENSITY)[:N, :, :N].copy()
        out[mixed] = (k, r)
        pm.close()
    ref = np.fft.rfftn(x)
    print("N=%d  r2c mixed vs numpy %.2e, transpose path vs numpy %.2e, roundtrip mixed %.2e" % (
        N, np.abs(out["1"][0] - ref).max() / np.abs(ref).max(), np.abs(out["0"][0] - ref).max() / np.abs(ref).max(),
        np.abs(out["1"][1] / N ** 3 - x).max()))
PY
for m in 1 0; do MGP_XFFT_MIXED=$m python tools/xfft_probe.py 320 5; done 2>&1 | tee gpurun_out/mx_probe_1gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for m in 1 0; do
  MGP_XFFT_MIXED=$m PROBE_ONLY=1,2,3,4,8 timeout 100 $TR tools/exchange_probe.py 320 8 2>&1 | grep -E "N=|rror"
  MGP_XFFT_MIXED=$m timeout 200 $TR bench.py --gpus 2 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/mx_bench2_320_mixed$m.json 2> gpurun_out/mx_bench2_320_mixed$m.err
  python -c "
import json; d=json.loads([x for x in open('gpurun_out/mx_bench2_320_mixed$m.json') if x.startswith('{')][-1]); print('mixed=$m ms/step %.3f' % d['ms_per_step'], d['roofline']['phases_ms'].get('FFT'))"
done 2>&1 | tee gpurun_out/mx_2gpu.txt
