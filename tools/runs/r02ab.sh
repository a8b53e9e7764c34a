# round 2, GPU call AB: work quantum 48 / 64 / 128 on the three configurations
mkdir -p gpurun_out
P=python
{
for q in 128 64 48 128 64; do
  echo "== LWFB_QUANTUM=$q"
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 127872 365 dav
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 10000 366 bea
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 100000 366 isobea
done
for q in 128 64; do
  echo "== LWFB_QUANTUM=$q, 1022976 members"
  LWFB_QUANTUM=$q LWFB_RECORD=0 LWFB_RUNS=1 $P tools/year_run.py 1022976 365 dav
done
} > gpurun_out/r02ab_quantum.log 2>&1
cat gpurun_out/r02ab_quantum.log
