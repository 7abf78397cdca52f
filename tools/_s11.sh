export TAG=r02m
K="snapshot or simple_pofk or readic or dropin_driver or baseline_sizes or power_spectrum" timeout 900 tools/gpu_session.sh newtests
