# round 2, GPU call M: hybrid callback-phase policy, 1-block-per-SM (255 register) build on the critical-path config, deep-profile timing
mkdir -p gpurun_out
P=python
{
export LWFB_QUANTUM=0
echo "== main (one launch)"
for cfg in dav isobea bea; do $P tools/profile_run.py 56832 20 $cfg; done
unset LWFB_QUANTUM
} > gpurun_out/r02m_prof.log 2>&1
{
echo "== main"
LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
for w in 5 8 12; do
  echo "== main LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
echo "== mb1 (one block per SM, 255 registers)"
LWFB_LIB=build_variants/mb1.so LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
LWFB_LIB=build_variants/mb1.so LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
for q in 64 256; do
  echo "== main LWFB_QUANTUM=$q"
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
} > gpurun_out/r02m_year.log 2>&1
{ $P tools/wide_run.py 4096 60 42; $P tools/wide_run.py 1184 30 100; $P tools/wide_run.py 296 10 400; } > gpurun_out/r02m_wide.log 2>&1
cat gpurun_out/r02m_prof.log gpurun_out/r02m_year.log gpurun_out/r02m_wide.log
