#!/bin/bash
# two GPUs: the multi-rank parity tests (world 2) on the fused path, then 2-GPU benches with and without it
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi -L | head -3
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu -k "2-" > gpurun_out/s4_mgpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s4_mgpu_tests.log; tail -6 gpurun_out/s4_mgpu_tests.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for N in 512 256; do for x in 1 0; do
  MGP_XFFT=$x timeout 400 $TR bench.py --gpus 2 --nmesh $N --steps 5 --warmup 3 > gpurun_out/s4_bench2_xf${x}_$N.json 2> gpurun_out/s4_bench2_xf${x}_$N.err
done; done
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s4_bench2_*.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, "ms/step %.3f" % d["ms_per_step"], {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm","MoveParticles","Sort")})
    except Exception as e: print(f, "failed", e)
PY
