export TAG=r02l
K="baseline_sizes or power_spectrum or rsd or nu_ or ic_download or test_ic or displacements_lcdm or mixed_radix" timeout 900 tools/gpu_session.sh newtests
BARGS="--steps 5 --warmup 3" BNAME=default tools/gpu_session.sh bench
