#!/bin/bash
# One parameterised GPU session script (run under gpurun): tools/gpu_session.sh <stage> [<stage> ...]
# Every stage is bounded by its own timeout and writes only under gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
TAG=${TAG:-r02}
NG=${NG:-1}
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --master-port ${PORT:-29533} --nproc-per-node"
for stage in "$@"; do
  echo "=== stage $stage ($(date +%T))"
  case $stage in
    experimental)      # first GPU run of the mixed-radix / wide-tile instances of the fused x-transform
      MGP_TEST_EXPERIMENTAL=1 timeout 300 python -m pytest tests/test_slab_fused.py -q -m gpu -k "mixed_radix_and_wide" -p no:cacheprovider \
        > gpurun_out/${TAG}_experimental.log 2>&1; tail -5 gpurun_out/${TAG}_experimental.log ;;
    tests)             # the whole single-GPU suite
      timeout 1500 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/${TAG}_gpu_tests.log 2>&1
      echo "rc=$?" >> gpurun_out/${TAG}_gpu_tests.log; tail -8 gpurun_out/${TAG}_gpu_tests.log ;;
    newtests)          # only the tests named in $K
      timeout 900 python -m pytest tests -x -q -m gpu -p no:cacheprovider -k "$K" > gpurun_out/${TAG}_newtests.log 2>&1
      echo "rc=$?" >> gpurun_out/${TAG}_newtests.log; tail -15 gpurun_out/${TAG}_newtests.log ;;
    bench)             # bench.py with $BARGS, record under the name $BNAME
      timeout 600 python bench.py $BARGS > gpurun_out/${TAG}_bench_${BNAME:-default}.json 2> gpurun_out/${TAG}_bench_${BNAME:-default}.err
      tail -c 3000 gpurun_out/${TAG}_bench_${BNAME:-default}.json; tail -3 gpurun_out/${TAG}_bench_${BNAME:-default}.err ;;
    benchmodes)        # deposit strategies at $NMESH (device step only)
      for m in ${MODES:-0 2}; do
        timeout 400 python bench.py --nmesh ${NMESH:-512} --steps 8 --warmup 4 --no-e2e --no-cpu-baseline --no-parity --deposit-mode $m $BARGS \
          > gpurun_out/${TAG}_bench_${NMESH:-512}_mode$m.json 2> gpurun_out/${TAG}_bench_${NMESH:-512}_mode$m.err
        python - <<PY
import json
try:
    d = json.loads([x for x in open("gpurun_out/${TAG}_bench_${NMESH:-512}_mode$m.json") if x.startswith("{")][-1])
    print("mode $m: ms/step %.3f" % d["ms_per_step"], d["roofline"]["phases_ms"])
except Exception as e:
    print("mode $m failed", e); print(open("gpurun_out/${TAG}_bench_${NMESH:-512}_mode$m.err").read()[-1500:])
PY
      done ;;
    launches)          # ncu launch list of one bench step (shares, not absolutes)
      timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c ${NCU_C:-400} --csv --log-file gpurun_out/${TAG}_launches.csv \
        python bench.py --nmesh ${NMESH:-512} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity ${BARGS} > gpurun_out/${TAG}_launches.log 2>&1
      tail -2 gpurun_out/${TAG}_launches.log ;;
    ncufull)           # ncu --set full of the kernels matching $KREGEX
      timeout 900 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX}" -s ${NCU_S:-0} -c ${NCU_C:-6} -f -o gpurun_out/${TAG}_${NCU_NAME:-prof} \
        python bench.py --nmesh ${NMESH:-512} --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-parity ${BARGS} > gpurun_out/${TAG}_ncufull.log 2>&1
      tail -2 gpurun_out/${TAG}_ncufull.log ;;
    mgpu_tests)        # multi-GPU parity on $NG GPUs, several tests at a time (tiny problems share the GPUs)
      timeout ${MGPU_TIMEOUT:-900} python -m pytest tests/test_multi_gpu.py -q -m gpu -p no:cacheprovider -n ${XDIST:-4} ${K:+-k "$K"} \
        > gpurun_out/${TAG}_mgpu_tests_${NG}gpu.log 2>&1
      echo "rc=$?" >> gpurun_out/${TAG}_mgpu_tests_${NG}gpu.log; tail -12 gpurun_out/${TAG}_mgpu_tests_${NG}gpu.log ;;
    probe)             # exchange pieces alone on $NG ranks
      for N in ${PROBE_N:-512 1024}; do
        PROBE_ONLY=${PROBE_ONLY:-0,1,2,3,4,8,9} timeout 150 $TR $NG tools/exchange_probe.py $N $NG 2>&1 | grep -E "N=|rror"
      done | tee gpurun_out/${TAG}_exchange_${NG}gpu.txt ;;
    mbench)            # bench.py on $NG ranks with $BARGS
      timeout 800 $TR $NG bench.py --gpus $NG $BARGS > gpurun_out/${TAG}_bench_${NG}gpu_${BNAME:-default}.json 2> gpurun_out/${TAG}_bench_${NG}gpu_${BNAME:-default}.err
      tail -c 2500 gpurun_out/${TAG}_bench_${NG}gpu_${BNAME:-default}.json; tail -3 gpurun_out/${TAG}_bench_${NG}gpu_${BNAME:-default}.err ;;
    pcie)              # pinned host <-> device copy rates on $NG ranks, alone and together
      timeout 200 $TR $NG tools/pcie_probe.py 2>&1 | grep -E "rank|aggregate|rror" | tee gpurun_out/${TAG}_pcie_${NG}gpu.txt ;;
    fof_lc)            # section 8(f).4: library tests, the two driver tests (their files kept), then the timing probe
      timeout 120 python -m pytest tests/test_lightcone.py tests/test_fof.py -q -m gpu -p no:cacheprovider -rA \
        --basetemp=gpurun_out/${TAG}_tmp_lib > gpurun_out/${TAG}_fof_lc_lib.log 2>&1
      echo "rc=$?" >> gpurun_out/${TAG}_fof_lc_lib.log; tail -6 gpurun_out/${TAG}_fof_lc_lib.log
      timeout 90 python -m pytest tests/test_dropin_driver.py -q -m gpu -p no:cacheprovider -rA -k "lightcone_driver or matchmaker_driver" \
        --basetemp=gpurun_out/${TAG}_tmp_drv > gpurun_out/${TAG}_fof_lc_drv.log 2>&1
      echo "rc=$?" >> gpurun_out/${TAG}_fof_lc_drv.log; tail -6 gpurun_out/${TAG}_fof_lc_drv.log
      find gpurun_out/${TAG}_tmp_lib gpurun_out/${TAG}_tmp_drv -type f -size +20M -delete 2>/dev/null
      timeout 70 python tools/fof_lc_probe.py > gpurun_out/${TAG}_fof_lc_probe.json 2> gpurun_out/${TAG}_fof_lc_probe.err
      tail -c 1500 gpurun_out/${TAG}_fof_lc_probe.json; tail -3 gpurun_out/${TAG}_fof_lc_probe.err
      timeout 50 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_fof_lc_launches.csv \
        python tools/fof_lc_probe.py --n1d 256 --lc-n1d 128 --reps 1 > gpurun_out/${TAG}_fof_lc_launches.log 2>&1
      tail -2 gpurun_out/${TAG}_fof_lc_launches.log ;;
    fof_ncu)           # ncu --set full of the halo finder's and the lightcone's own kernels
      timeout 90 ncu --set full --clock-control none --import-source on -k "regex:LinkStep|PropsStep|k_lightcone" -c 7 -f -o gpurun_out/${TAG}_fof_lc \
        python tools/fof_lc_probe.py --reps 1 > gpurun_out/${TAG}_fof_ncu.log 2>&1
      tail -2 gpurun_out/${TAG}_fof_ncu.log ;;
    smi)
      nvidia-smi --query-gpu=index,name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv | tee gpurun_out/${TAG}_smi.txt
      nvidia-smi topo -m | head -12 | tee -a gpurun_out/${TAG}_smi.txt; nproc | tee -a gpurun_out/${TAG}_smi.txt; free -g | head -2 | tee -a gpurun_out/${TAG}_smi.txt ;;
    *) echo "unknown stage $stage" ;;
  esac
done
echo "=== done ($(date +%T))"
