#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533"
for N in 512 256; do timeout 300 $TR tools/exchange_probe.py $N 8 2>&1 | grep "N=" ; done | tee gpurun_out/s5_exchange_2gpu.txt
timeout 300 $TR tools/exchange_probe.py 512 4 2>&1 | grep "N=" | tee -a gpurun_out/s5_exchange_2gpu.txt
python tools/exchange_probe.py 512 8 2>&1 | grep "N=" | tee gpurun_out/s5_exchange_1gpu.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -x -q -m gpu > gpurun_out/s5_mgpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s5_mgpu_tests.log; tail -6 gpurun_out/s5_mgpu_tests.log
