# round 2, GPU call C: warp-level callback phases (counted stage barriers) + cheaper exp tail
mkdir -p gpurun_out
P=python
timeout 1500 $P -m pytest tests -m gpu -x -q > gpurun_out/r02c_pytest.log 2>&1; echo "pytest rc $?" >> gpurun_out/r02c_pytest.log
{
echo "== legacy kernel diagnostics, current build"; $P tools/legacy_diag.py
echo "== legacy kernel diagnostics, NVVM -O1 build"; LWFB_LIB=build_variants/o1.so $P tools/legacy_diag.py
} > gpurun_out/r02c_legacy.log 2>&1
export LWFB_QUANTUM=0
{
for cfg in dav davnoiso bea isobea; do $P tools/profile_run.py 56832 20 $cfg; done
} > gpurun_out/r02c_prof.log 2>&1
unset LWFB_QUANTUM
$P tools/year_run.py 127872 365 > gpurun_out/r02c_year.log 2>&1
{
$P tools/profile_run.py 10000 366 bea
$P tools/profile_run.py 100000 366 isobea
} > gpurun_out/r02c_cfg.log 2>&1
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02c_lanes $P tools/profile_run.py 56832 20 dav > gpurun_out/r02c_ncu.log 2>&1
tail -4 gpurun_out/r02c_pytest.log; cat gpurun_out/r02c_legacy.log gpurun_out/r02c_prof.log gpurun_out/r02c_year.log gpurun_out/r02c_cfg.log
