# round 2, GPU call F: lean / fixed-N instantiations, per-member shared slots + cp.async forcing staging
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/lean_nx0.so build_variants/lean.so build_variants/stage.so build_variants/stage2.so; do
  echo "== $lib"
  for cfg in dav isobea; do LWFB_LIB=$lib $P tools/profile_run.py 56832 20 $cfg; done
done
} > gpurun_out/r02f_prof.log 2>&1
unset LWFB_QUANTUM
LWFB_LIB=build_variants/stage2.so timeout 1500 $P -m pytest tests -m gpu -x -q > gpurun_out/r02f_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02f_pytest.log
LWFB_LIB=build_variants/stage2.so $P tools/year_run.py 127872 365 > gpurun_out/r02f_year.log 2>&1
cat gpurun_out/r02f_prof.log gpurun_out/r02f_year.log; tail -3 gpurun_out/r02f_pytest.log
