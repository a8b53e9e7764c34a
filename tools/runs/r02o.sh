# round 2, GPU call O (surplus lanes fixed): callback core in one lane per member, latency (255 register) build for launches that do not fill the
# device, hybrid callback-phase policy: whole GPU suite, timings, ncu --set full of the lane kernel
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02o_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02o_pytest.log
{
export LWFB_QUANTUM=0
echo "== main (one launch)"
for cfg in dav isobea bea; do $P tools/profile_run.py 56832 20 $cfg; done
unset LWFB_QUANTUM
} > gpurun_out/r02o_prof.log 2>&1
{
echo "== main"
LWFB_RECORD=0 LWFB_RUNS=2 $P tools/year_run.py 127872 365 dav
LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
LWFB_TRACE=gpurun_out/r02o_trace_bea10k.txt LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
echo "== main, LWFB_LATENCY_KERNEL=0"
LWFB_LATENCY_KERNEL=0 LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
LWFB_LATENCY_KERNEL=0 LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
} > gpurun_out/r02o_year.log 2>&1
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02o_lanes $P tools/profile_run.py 56832 20 dav > gpurun_out/r02o_ncu.log 2>&1
unset LWFB_QUANTUM
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02o_pytest.log; tail -4 gpurun_out/r02o_pytest.log; cat gpurun_out/r02o_prof.log gpurun_out/r02o_year.log; tail -3 gpurun_out/r02o_ncu.log
