# round 2, GPU call X (2 GPUs): the bench under torchrun (strong scaling of the 1M ensemble, NCCL all-gather of the scores),
# the multi-GPU test of the C host on two real devices
mkdir -p gpurun_out
P=python
$P -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2 --warmup 1 --daily-members 0 > gpurun_out/r02x_bench_2gpu.json 2> gpurun_out/r02x_bench_2gpu.err
timeout 900 $P -m pytest tests -m gpu -x -q -k "c_host or multi_handle" > gpurun_out/r02x_pytest_multi.log 2>&1
tail -c 1200 gpurun_out/r02x_bench_2gpu.json; tail -3 gpurun_out/r02x_bench_2gpu.err; tail -3 gpurun_out/r02x_pytest_multi.log
