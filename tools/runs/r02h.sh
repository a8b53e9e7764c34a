# round 2, GPU call H: integer-domain log reduction (128-entry table)
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
{
for cfg in dav isobea bea; do $P tools/profile_run.py 56832 20 $cfg; done
} > gpurun_out/r02h_prof.log 2>&1
unset LWFB_QUANTUM
timeout 1500 $P -m pytest tests -m gpu -x -q > gpurun_out/r02h_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02h_pytest.log
$P tools/year_run.py 127872 365 > gpurun_out/r02h_year.log 2>&1
cat gpurun_out/r02h_prof.log gpurun_out/r02h_year.log; tail -3 gpurun_out/r02h_pytest.log
