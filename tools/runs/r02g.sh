# round 2, GPU call G: shared-memory exchange of the ITER minimum and the accumulator rates; fixed-N on/off
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/cur_nx0.so; do
  echo "== $lib"
  for cfg in dav isobea bea; do LWFB_LIB=$lib $P tools/profile_run.py 56832 20 $cfg; done
done
} > gpurun_out/r02g_prof.log 2>&1
unset LWFB_QUANTUM
timeout 1500 $P -m pytest tests -m gpu -x -q > gpurun_out/r02g_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02g_pytest.log
$P tools/year_run.py 127872 365 > gpurun_out/r02g_year.log 2>&1
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02g_lanes $P tools/profile_run.py 56832 20 dav > gpurun_out/r02g_ncu.log 2>&1
cat gpurun_out/r02g_prof.log gpurun_out/r02g_year.log; tail -3 gpurun_out/r02g_pytest.log
