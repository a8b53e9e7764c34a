# round 2, GPU call AA: parameter sweeps of the work-quantum execution after the callback phase became cheaper (same box),
# isotope species loop unrolled (variant)
mkdir -p gpurun_out
P=python
{
echo "== main"
LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
echo "== adeunroll"
LWFB_LIB=build_variants/adeunroll.so LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
LWFB_LIB=build_variants/adeunroll.so LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
for w in 2 3 8; do
  echo "== main LWFB_CB_WAIT=$w"
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_CB_WAIT=$w LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
for q in 64 96 192; do
  echo "== main LWFB_QUANTUM=$q"
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
for pp in 8 25; do
  echo "== main LWFB_PRIO_PCT=$pp"
  LWFB_PRIO_PCT=$pp LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_PRIO_PCT=$pp LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
echo "== main (again)"
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
} > gpurun_out/r02aa_sweep.log 2>&1
cat gpurun_out/r02aa_sweep.log
