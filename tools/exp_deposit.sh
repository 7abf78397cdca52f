#!/bin/bash
# developer experiment: deposit variants early (warm-up 3) and late (warm-up 26) in the run
for w in 3 26; do for agg in 1 0; do for mode in 0 1; do
  MGP_AGG=$agg timeout 300 python bench.py --steps 3 --warmup $w --no-e2e --no-cpu-baseline --no-parity --deposit-mode $mode --sort-interval 1 2>/dev/null \
   | python -c "import json,sys; d=json.loads(sys.stdin.read()); p=d['roofline']['phases_ms']; print('warmup=$w agg=$agg mode=$mode', round(d['ms_per_step'],3), 'PtoMesh', p['PtoMesh'], 'MtoP', p['MtoParticles'], 'Sort', p.get('Sort'))"
done; done; done
