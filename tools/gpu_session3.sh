#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_slab_fused.py -x -q -m gpu --durations=5 > gpurun_out/s3_slab_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s3_slab_tests.log; tail -12 gpurun_out/s3_slab_tests.log
for N in 256 512; do for x in 1 0; do MGP_XFFT=$x python tools/xfft_probe.py $N 5; done; done 2>&1 | tee gpurun_out/s3_probe.txt
MGP_XFFT=1 python tools/xfft_probe.py 512 5 4 2>&1 | tee -a gpurun_out/s3_probe.txt
MGP_XFFT=1 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s3_launches_xf1.csv python tools/xfft_probe.py 512 1 > /dev/null 2>&1
grep -E "k_xfft" gpurun_out/s3_launches_xf1.csv | awk -F'","' '{print substr($5,1,60), $(NF)}' | head
ncu --set full --clock-control none --import-source on -k regex:k_xfft_bwd -c 1 -f -o gpurun_out/s3_xfft_bwd python tools/xfft_probe.py 512 1 > gpurun_out/s3_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_xfft_fwd -c 1 -f -o gpurun_out/s3_xfft_fwd python tools/xfft_probe.py 512 1 >> gpurun_out/s3_ncu.log 2>&1
B="python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline"
MGP_FORCE_SLAB=1 MGP_XFFT=1 timeout 300 $B --nmesh 512 > gpurun_out/s3_bench_slab_xf1_512.json 2> gpurun_out/s3_bench_slab_xf1_512.err
python - <<'PY'
import json
try:
    d=json.loads([x for x in open("gpurun_out/s3_bench_slab_xf1_512.json") if x.startswith('{')][-1])
    print("slab xf1 512: ms/step %.3f" % d["ms_per_step"], d["roofline"]["phases_ms"])
except Exception as e: print("failed", e)
PY
