# round 2, GPU call Y: GWAT bookkeeping of the lean kernels in shared slots (one lane per member) against the previous build
mkdir -p gpurun_out
P=python
{
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/divz.so; do
  echo "== $lib (rep $rep)"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
done
export LWFB_QUANTUM=0
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/divz.so; do
  echo "== $lib (one launch)"
  for cfg in dav isobea; do LWFB_LIB=$lib $P tools/profile_run.py 56832 20 $cfg; done
done
} > gpurun_out/r02y_year.log 2>&1
cat gpurun_out/r02y_year.log
