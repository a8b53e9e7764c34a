# round 2, GPU call Z: validation of the shipped build - whole GPU suite, bench lines of the three BASELINE configurations,
# ncu --set full of the lane kernel (DAV2020, isoBEA2016 profile workloads)
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02z_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02z_pytest.log
$P -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02z_smoke.log 2>&1
$P bench.py > gpurun_out/r02z_bench_dav2020.json 2> gpurun_out/r02z_bench_dav2020.err
$P bench.py --config isobea2016 --steps 2 --warmup 3 > gpurun_out/r02z_bench_isobea2016.json 2> gpurun_out/r02z_bench_isobea2016.err
$P bench.py --config bea2016 --steps 2 --warmup 3 > gpurun_out/r02z_bench_bea2016.json 2> gpurun_out/r02z_bench_bea2016.err
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02z_lanes_dav $P tools/profile_run.py 56832 20 dav > gpurun_out/r02z_ncu_dav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02z_lanes_isobea $P tools/profile_run.py 28416 10 isobea > gpurun_out/r02z_ncu_isobea.log 2>&1
unset LWFB_QUANTUM
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02z_pytest.log; tail -4 gpurun_out/r02z_pytest.log; cat gpurun_out/r02z_smoke.log | tail -2
for f in dav2020 isobea2016 bea2016; do head -c 160 gpurun_out/r02z_bench_$f.json; echo; tail -2 gpurun_out/r02z_bench_$f.err; done
