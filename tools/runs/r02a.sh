# round 2, GPU call A: fresh baseline of the r01 build + free-running / thin-front discovery
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
{
for cfg in dav davnoiso; do
  $P tools/profile_run.py 56832 20 $cfg
  LWFB_LIB=build_variants/free.so $P tools/profile_run.py 56832 20 $cfg
done
} > gpurun_out/r02a_free.log 2>&1
unset LWFB_QUANTUM
{
for tf in 0 1 2; do LWFB_THIN_FRONT=$tf $P tools/profile_run.py 10000 366 bea; done
LWFB_LIB=build_variants/free.so $P tools/profile_run.py 10000 366 bea
} > gpurun_out/r02a_bea.log 2>&1
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02a_lanes_v14 $P tools/profile_run.py 56832 20 dav > gpurun_out/r02a_ncu.log 2>&1
LWFB_LIB=build_variants/free.so ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02a_lanes_free_noiso $P tools/profile_run.py 56832 20 davnoiso > gpurun_out/r02a_ncu2.log 2>&1
cat gpurun_out/r02a_free.log gpurun_out/r02a_bea.log
tail -3 gpurun_out/r02a_ncu.log
