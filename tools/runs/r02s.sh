# round 2, GPU call S: why is a single launch with a 2-pass waiting bound (LWFB_CB_WAIT=2) pathological with the block-level callback?
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/chain.so; do
for w in 2 4 32; do
  echo "== $lib LWFB_CB_WAIT=$w"
  LWFB_LIB=$lib LWFB_CB_WAIT=$w $P tools/profile_run.py 14208 5 dav
done
done
} > gpurun_out/r02s_small.log 2>&1
LWFB_CB_WAIT=2 ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02s_wait2 $P tools/profile_run.py 14208 5 dav > gpurun_out/r02s_ncu.log 2>&1
cat gpurun_out/r02s_small.log; tail -3 gpurun_out/r02s_ncu.log
