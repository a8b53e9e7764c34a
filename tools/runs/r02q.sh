# round 2, GPU call Q: same-box A/B of the INFLOW sweeps against the bottom-up chain; cost of a callback phase
mkdir -p gpurun_out
P=python
{
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/chain.so; do
  echo "== $lib (rep $rep)"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
done
done
} > gpurun_out/r02q_year.log 2>&1
{
export LWFB_QUANTUM=0
for w in 2 8 32; do
  echo "== one launch, LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w $P tools/profile_run.py 56832 20 dav
done
unset LWFB_QUANTUM
} > gpurun_out/r02q_cbcost.log 2>&1
cat gpurun_out/r02q_year.log gpurun_out/r02q_cbcost.log
