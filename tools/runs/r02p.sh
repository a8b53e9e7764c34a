# round 2, GPU call P: INFLOW limiter as converging parallel sweeps; both builds of the lane kernel in the parity test
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02p_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02p_pytest.log
{
export LWFB_QUANTUM=0
echo "== main (one launch)"
for cfg in dav isobea bea; do $P tools/profile_run.py 56832 20 $cfg; done
unset LWFB_QUANTUM
} > gpurun_out/r02p_prof.log 2>&1
{
echo "== main"
LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
LWFB_TRACE=gpurun_out/r02p_trace_isobea.txt LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
} > gpurun_out/r02p_year.log 2>&1
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02p_pytest.log; tail -4 gpurun_out/r02p_pytest.log; cat gpurun_out/r02p_prof.log gpurun_out/r02p_year.log
