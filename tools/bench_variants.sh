#!/bin/bash
# developer tool: bench.py over deposit mode x sort interval (x grid precision); prints ms/step + phases
for cfg in "0 1 8" "0 4 8" "0 0 8" "2 1 8" "0 4 4"; do
  set -- $cfg
  timeout 300 python bench.py --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-parity --scale-dependent 0 --deposit-mode $1 --sort-interval $2 --grid-bytes $3 2>gpurun_out/variant_err.log \
   | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('mode=$1 sort=$2 g=$3', round(d['ms_per_step'],3), d['roofline']['phases_ms'])" || tail -5 gpurun_out/variant_err.log
done
