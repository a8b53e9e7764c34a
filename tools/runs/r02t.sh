# round 2, GPU call T: bench lines of the three BASELINE configurations with the build of call R, the reference arm,
# reproducibility of the single-launch LWFB_CB_WAIT=2 anomaly
mkdir -p gpurun_out
P=python
$P bench.py > gpurun_out/r02t_bench_dav2020.json 2> gpurun_out/r02t_bench_dav2020.err
$P bench.py --config isobea2016 --steps 2 --warmup 3 > gpurun_out/r02t_bench_isobea2016.json 2> gpurun_out/r02t_bench_isobea2016.err
$P bench.py --config bea2016 --steps 2 --warmup 3 > gpurun_out/r02t_bench_bea2016.json 2> gpurun_out/r02t_bench_bea2016.err
$P bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02t_bench_reference.json 2> gpurun_out/r02t_bench_reference.err
{
export LWFB_QUANTUM=0
for rep in 1 2; do
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/chain.so; do
  echo "== $lib LWFB_CB_WAIT=2 (rep $rep)"
  LWFB_LIB=$lib LWFB_CB_WAIT=2 $P tools/profile_run.py 56832 20 dav
done
done
LWFB_CB_WAIT=2 $P tools/profile_run.py 56832 10 dav
LWFB_CB_WAIT=2 $P tools/profile_run.py 28416 20 dav
unset LWFB_QUANTUM
} > gpurun_out/r02t_wait2.log 2>&1
for f in dav2020 isobea2016 bea2016 reference; do tail -c 700 gpurun_out/r02t_bench_$f.json; echo; tail -2 gpurun_out/r02t_bench_$f.err; done
cat gpurun_out/r02t_wait2.log
