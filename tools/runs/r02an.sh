# round 2, GPU call AN (2 GPUs): bench.py --balance 10 on isoBEA2016 (the --balance 0 line is from call AM)
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --config isobea2016 --balance 10 --steps 1 --warmup 1 --e2e-steps 1 --daily-members 0 > gpurun_out/r02am_bench_isobea_2gpu_balance10.json 2> gpurun_out/r02am_balance10.err
grep -v "^\*\|OMP_NUM" gpurun_out/r02am_balance10.err | tail -5
