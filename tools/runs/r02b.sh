# round 2, GPU call B: member-major layouts + in-run scores: GPU parity suite, profile workload, no-isotope anomaly
mkdir -p gpurun_out
P=python
timeout 1500 $P -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02b_pytest.log
export LWFB_QUANTUM=0
{
$P tools/profile_run.py 56832 20 dav
$P tools/profile_run.py 56832 20 davnoiso
$P tools/profile_run.py 56832 20 davnoiso
LWFB_RECORD=0 $P tools/profile_run.py 56832 20 davnoiso
$P tools/profile_run.py 56832 20 bea
$P tools/profile_run.py 56832 20 isobea
} > gpurun_out/r02b_prof.log 2>&1
unset LWFB_QUANTUM
$P tools/year_run.py 127872 365 > gpurun_out/r02b_year.log 2>&1
tail -5 gpurun_out/r02b_pytest.log; cat gpurun_out/r02b_prof.log gpurun_out/r02b_year.log
