# round 2, GPU call U: ncu --set full of the shipped lane kernel on the DAV2020 and isoBEA2016 profile workloads
mkdir -p gpurun_out
P=python
export LWFB_QUANTUM=0
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02u_lanes_dav $P tools/profile_run.py 56832 20 dav > gpurun_out/r02u_ncu_dav.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:integrate_lanes -c 1 -f -o gpurun_out/r02u_lanes_isobea $P tools/profile_run.py 28416 10 isobea > gpurun_out/r02u_ncu_isobea.log 2>&1
unset LWFB_QUANTUM
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02u_launch_list.csv $P bench.py --members 127872 --steps 2 --warmup 1 --no-cpu-baseline --no-parity --daily-members 0 > gpurun_out/r02u_launch_list.log 2>&1
tail -2 gpurun_out/r02u_ncu_dav.log gpurun_out/r02u_ncu_isobea.log; tail -c 300 gpurun_out/r02u_launch_list.log; wc -l gpurun_out/r02u_launch_list.csv
