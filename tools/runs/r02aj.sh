# round 2, GPU call AJ: 512-thread blocks (one lock-step group of sixteen warps per SM) against the shipped 256-thread blocks
mkdir -p gpurun_out
P=python
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/blk512.so lwfbrook90.jl_b200/liblwfb90.so build_variants/blk512.so; do
  echo "== $lib"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
} > gpurun_out/r02ai_blk512.log 2>&1
cat gpurun_out/r02ai_blk512.log
