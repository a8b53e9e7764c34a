# round 2, GPU call AD (4 GPUs): the bench under torchrun, strong scaling of the 1M ensemble
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 2 --warmup 1 --daily-members 0 > gpurun_out/r02ad_bench_4gpu.json 2> gpurun_out/r02ad_bench_4gpu.err
head -c 400 gpurun_out/r02ad_bench_4gpu.json; tail -3 gpurun_out/r02ad_bench_4gpu.err
