# round 2, GPU call R: block-level daily callback (one member per lane) against the previous build (chain.so = commit 2e of
# the same day, callback in one lane per member of every warp) on the same box; whole GPU suite incl. the C host test
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02r_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02r_pytest.log
{
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/chain.so; do
  echo "== $lib (rep $rep)"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
done
done
for w in 3 8; do
  echo "== main LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
} > gpurun_out/r02r_year.log 2>&1
{
export LWFB_QUANTUM=0
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/chain.so; do
  echo "== $lib (one launch)"
  for cfg in dav isobea bea; do LWFB_LIB=$lib $P tools/profile_run.py 56832 20 $cfg; done
done
for w in 2 8 32; do
  echo "== main, one launch, LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w $P tools/profile_run.py 56832 20 dav
done
unset LWFB_QUANTUM
} > gpurun_out/r02r_prof.log 2>&1
grep -a "tolerance vs libm\|setup_members\]\|multi_host" gpurun_out/r02r_pytest.log; tail -4 gpurun_out/r02r_pytest.log; cat gpurun_out/r02r_year.log gpurun_out/r02r_prof.log
