# round 2, GPU call AC (work quantum 64): validation of the shipped build - whole GPU suite, bench lines of the three BASELINE configurations,
mkdir -p gpurun_out
P=python
timeout 2400 $P -m pytest tests -m gpu -x -q -s > gpurun_out/r02ac_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02ac_pytest.log
$P -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r02ac_smoke.log 2>&1
$P bench.py > gpurun_out/r02ac_bench_dav2020.json 2> gpurun_out/r02ac_bench_dav2020.err
$P bench.py --config isobea2016 --steps 2 --warmup 3 > gpurun_out/r02ac_bench_isobea2016.json 2> gpurun_out/r02ac_bench_isobea2016.err
$P bench.py --config bea2016 --steps 2 --warmup 3 > gpurun_out/r02ac_bench_bea2016.json 2> gpurun_out/r02ac_bench_bea2016.err
grep -a "tolerance vs libm\|setup_members\]" gpurun_out/r02ac_pytest.log; tail -4 gpurun_out/r02ac_pytest.log; cat gpurun_out/r02ac_smoke.log | tail -2
for f in dav2020 isobea2016 bea2016; do head -c 160 gpurun_out/r02ac_bench_$f.json; echo; tail -2 gpurun_out/r02ac_bench_$f.err; done
