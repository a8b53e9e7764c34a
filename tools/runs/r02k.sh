# round 2, GPU call K: BASELINE configs 3 and 4 through bench.py (the same line as the headline), thin front blocks measured
mkdir -p gpurun_out
P=python
$P bench.py --config bea2016 --steps 2 --warmup 3 > gpurun_out/r02k_bench_bea2016.json 2> gpurun_out/r02k_bench_bea2016.err
$P bench.py --config isobea2016 --steps 2 --warmup 3 > gpurun_out/r02k_bench_isobea2016.json 2> gpurun_out/r02k_bench_isobea2016.err
for tf in 1 2; do
LWFB_THIN_FRONT=$tf $P bench.py --config bea2016 --steps 2 --warmup 3 --no-cpu-baseline --no-parity --daily-members 0 > gpurun_out/r02k_bench_bea2016_thinfront$tf.json 2>&1
done
tail -c 1500 gpurun_out/r02k_bench_bea2016.json; tail -c 1500 gpurun_out/r02k_bench_isobea2016.json; tail -c 600 gpurun_out/r02k_bench_bea2016_thinfront*.json; tail -3 gpurun_out/*.err
