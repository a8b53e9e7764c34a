# round 2, GPU call V: INFLOW chain trimmed to the range of limiting layers, against the build of call R (blockcb.so), same box
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q > gpurun_out/r02v_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02v_pytest.log
{
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/blockcb.so; do
  echo "== $lib (rep $rep)"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
done
done
} > gpurun_out/r02v_year.log 2>&1
tail -4 gpurun_out/r02v_pytest.log; cat gpurun_out/r02v_year.log
