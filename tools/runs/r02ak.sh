# round 2, GPU call AK: limiter test decided by a monotone bound (no division per lane and evaluation) against the previous build
mkdir -p gpurun_out
P=python
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/prev.so lwfbrook90.jl_b200/liblwfb90.so build_variants/prev.so; do
  echo "== $lib"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
} > gpurun_out/r02ak_bound.log 2>&1
cat gpurun_out/r02ak_bound.log
