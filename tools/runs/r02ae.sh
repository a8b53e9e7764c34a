# round 2, GPU call AE (8 GPUs): the bench under torchrun, strong scaling of the 1M ensemble
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 2 --warmup 1 --daily-members 0 > gpurun_out/r02ad_bench_8gpu.json 2> gpurun_out/r02ad_bench_8gpu.err
head -c 400 gpurun_out/r02ad_bench_8gpu.json; tail -3 gpurun_out/r02ad_bench_8gpu.err
