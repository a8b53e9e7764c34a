# round 2, GPU call L: table exp + Estrin log + waste-weighted callback phases + warp-per-member kernel (N > 32)
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02l_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02l_pytest.log
{
export LWFB_QUANTUM=0
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/waitpol.so; do
  echo "== $lib (one launch)"
  for cfg in dav isobea bea; do LWFB_LIB=$lib $P tools/profile_run.py 56832 20 $cfg; done
done
for w in 2 5 8; do
  echo "== main, LWFB_CB_WAIT=$w (one launch)"
  for cfg in dav isobea; do LWFB_CB_WAIT=$w $P tools/profile_run.py 56832 20 $cfg; done
done
unset LWFB_QUANTUM
} > gpurun_out/r02l_prof.log 2>&1
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/waitpol.so; do
  echo "== $lib"
  LWFB_LIB=$lib LWFB_RECORD=0 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 100000 366 isobea
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 10000 366 bea
done
for w in 2 5; do
  echo "== main LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 100000 366 isobea
done
LWFB_TRACE=gpurun_out/r02l_trace_isobea.txt LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
LWFB_TRACE=gpurun_out/r02l_trace_bea10k.txt LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
} > gpurun_out/r02l_year.log 2>&1
$P tools/wide_run.py 4096 60 > gpurun_out/r02l_wide.log 2>&1
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02l_pytest.log; tail -5 gpurun_out/r02l_pytest.log; cat gpurun_out/r02l_prof.log gpurun_out/r02l_year.log gpurun_out/r02l_wide.log
