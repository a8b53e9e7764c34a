# round 2, GPU call AF: fewer stage barriers (LWFB_STAGE_SYNC_MASK) on DAV2020 and isoBEA2016, same box
mkdir -p gpurun_out
P=python
{
for lib in lwfbrook90.jl_b200/liblwfb90.so build_variants/sync135.so build_variants/sync14.so build_variants/sync1.so lwfbrook90.jl_b200/liblwfb90.so; do
  echo "== $lib"
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_LIB=$lib LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
} > gpurun_out/r02af_stagesync.log 2>&1
cat gpurun_out/r02af_stagesync.log
