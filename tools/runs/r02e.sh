# round 2, GPU call E: div_c arithmetic, device-side setup, new tests
mkdir -p gpurun_out
P=python
timeout 1800 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02e_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02e_pytest.log
export LWFB_QUANTUM=0
{
for cfg in dav davnoiso bea isobea; do $P tools/profile_run.py 56832 20 $cfg; done
} > gpurun_out/r02e_prof.log 2>&1
unset LWFB_QUANTUM
$P tools/year_run.py 127872 365 > gpurun_out/r02e_year.log 2>&1
$P bench.py --steps 1 --warmup 1 --total-members 127872 --e2e-steps 1 > gpurun_out/r02e_bench_small.json 2> gpurun_out/r02e_bench_small.err
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02e_lanes $P tools/profile_run.py 56832 20 dav > gpurun_out/r02e_ncu.log 2>&1
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02e_pytest.log; tail -4 gpurun_out/r02e_pytest.log; cat gpurun_out/r02e_prof.log gpurun_out/r02e_year.log; tail -3 gpurun_out/r02e_bench_small.err; cat gpurun_out/r02e_bench_small.json
