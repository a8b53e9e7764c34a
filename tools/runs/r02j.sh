# round 2, GPU call J: validation of the shipped build + the metric's own configuration (1,022,976 members on one GPU)
mkdir -p gpurun_out
P=python
timeout 1800 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02j_pytest.log
export LWFB_QUANTUM=0
{
for cfg in dav isobea bea; do $P tools/profile_run.py 56832 20 $cfg; done
} > gpurun_out/r02j_prof.log 2>&1
unset LWFB_QUANTUM
$P tools/year_run.py 127872 365 > gpurun_out/r02j_year.log 2>&1
$P bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02j_lanes $P tools/profile_run.py 56832 20 dav > gpurun_out/r02j_ncu.log 2>&1
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02j_pytest.log; tail -4 gpurun_out/r02j_pytest.log; cat gpurun_out/r02j_prof.log gpurun_out/r02j_year.log; tail -3 gpurun_out/r02j_bench.err; cat gpurun_out/r02j_bench.json
