export TAG=r02j NG=8
B="--nmesh 1024 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline"
BARGS="$B" BNAME=1024_default tools/gpu_session.sh mbench
MGP_XFFT_WIDE=1 BARGS="$B" BNAME=1024_wide tools/gpu_session.sh mbench
MGP_XFFT_DMA=1 BARGS="$B" BNAME=1024_dma tools/gpu_session.sh mbench
MGP_XFFT_CPS=2 BARGS="$B" BNAME=1024_cps2 tools/gpu_session.sh mbench
MGP_XFFT_WIDE=1 PROBE_N="1024" PROBE_ONLY=1,2,8 TAG=r02j_wide tools/gpu_session.sh probe
MGP_XFFT_DMA=1 PROBE_N="1024" PROBE_ONLY=1,2,6,7,8 TAG=r02j_dma tools/gpu_session.sh probe
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/r02j_bench_8gpu_*.json")):
    try:
        d=json.loads([x for x in open(f) if x.startswith('{')][-1])
        print(f, "ms/step %.3f  step frac %.3f" % (d["ms_per_step"], d["roofline"]["step"]["frac"]), {k:v for k,v in d["roofline"]["phases_ms"].items() if k in ("FFT","Comm","MoveParticles","Sort","PtoMesh","MtoParticles")})
    except Exception as e: print(f, "failed", e)
PY
