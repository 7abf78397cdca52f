export TAG=r02o NG=8
BARGS="--steps 10 --warmup 3 --no-cpu-baseline" BNAME=final tools/gpu_session.sh mbench
