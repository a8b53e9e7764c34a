# round 2, GPU call W: saturated layer's room term without the division slow path (div_z), against the build of call R
mkdir -p gpurun_out
P=python
{
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/blockcb.so; do
  echo "== $lib (rep $rep)"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
done
done
} > gpurun_out/r02w_year.log 2>&1
cat gpurun_out/r02w_year.log
