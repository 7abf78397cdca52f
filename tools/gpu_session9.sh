#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 400 python -m pytest tests -x -q -m gpu > gpurun_out/s9_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/s9_gpu_tests.log; tail -4 gpurun_out/s9_gpu_tests.log
timeout 200 python bench.py > gpurun_out/s9_bench.json 2> gpurun_out/s9_bench.err; cut -c1-400 gpurun_out/s9_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
