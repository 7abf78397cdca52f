export NG=8 MGP_XFFT_WIDE=1 PROBE_N="1024" PROBE_ONLY=1,8,9
for tr in 1 38 75; do
  MGP_XFFT_TRIM=$tr TAG=r02k_trim$tr tools/gpu_session.sh probe
done
MGP_XFFT_TRIM=1 MGP_XFFT_CPS=2 TAG=r02k_cps2 tools/gpu_session.sh probe
