#!/bin/bash
# one GPU: new slab-path tests, slab probes (fused vs cuFFT+transpose vs 3-D plans), default bench, then the full GPU suite
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_smi.txt 2>&1
echo "== slab tests" ; timeout 600 python -m pytest tests/test_slab_fused.py -x -q -m gpu --durations=8 > gpurun_out/s1_slab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s1_slab_tests.log; tail -5 gpurun_out/s1_slab_tests.log
B="python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline"
echo "== probes"
for N in 256 512; do
  MGP_FORCE_SLAB=1 MGP_XFFT=1 timeout 300 $B --nmesh $N > gpurun_out/s1_probe_slab_xf1_$N.json 2> gpurun_out/s1_probe_slab_xf1_$N.err
  MGP_FORCE_SLAB=1 MGP_XFFT=0 timeout 300 $B --nmesh $N > gpurun_out/s1_probe_slab_xf0_$N.json 2> gpurun_out/s1_probe_slab_xf0_$N.err
  timeout 300 $B --nmesh $N > gpurun_out/s1_probe_3d_$N.json 2> gpurun_out/s1_probe_3d_$N.err
done
MGP_FORCE_SLAB=1 MGP_XFFT=1 MGP_XFFT_TK=4 timeout 300 $B --nmesh 512 > gpurun_out/s1_probe_slab_xf1_tk4_512.json 2>&1
MGP_FORCE_SLAB=1 MGP_XFFT=1 MGP_XFFT_TK=16 timeout 300 $B --nmesh 512 > gpurun_out/s1_probe_slab_xf1_tk16_512.json 2>&1
for f in gpurun_out/s1_probe_*.json; do echo $f; python - "$f" <<'PY'
import json,sys
try:
    l=[x for x in open(sys.argv[1]) if x.startswith('{')][-1]; d=json.loads(l)
    print("  ms/step %.3f" % d["ms_per_step"], {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm","Forces","SDField")})
except Exception as e: print("  failed", e)
PY
done
echo "== bench" ; timeout 600 python bench.py > gpurun_out/s1_bench.json 2> gpurun_out/s1_bench.err; tail -c 600 gpurun_out/s1_bench.json
echo "== full gpu suite"; timeout 1500 python -m pytest tests -x -q -m gpu --durations=15 --deselect tests/test_slab_fused.py > gpurun_out/s1_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s1_gpu_tests.log; tail -25 gpurun_out/s1_gpu_tests.log
