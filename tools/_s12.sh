export TAG=r02n NG=2
K="(2- or [2] or 2-0 or 2-3 or False-2 or True-2)" XDIST=3 MGPU_TIMEOUT=600 tools/gpu_session.sh mgpu_tests
timeout 600 python -m pytest tests/test_dropin_driver.py -q -m gpu -p no:cacheprovider -k "on_ranks" > gpurun_out/r02n_driver_ranks.log 2>&1; tail -5 gpurun_out/r02n_driver_ranks.log
BARGS="--nmesh 256 --target-gpus 2 --target-nmesh 512 --steps 4 --warmup 3" BNAME=flowtest tools/gpu_session.sh mbench
