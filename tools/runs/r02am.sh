# round 2, GPU call AM (2 GPUs): cost-balanced shards in the one-process-per-GPU path (bench.py --balance) on isoBEA2016
mkdir -p gpurun_out
for b in 10 0; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2953$b bench.py --gpus 2 --config isobea2016 --balance $b --steps 1 --warmup 1 --e2e-steps 1 --daily-members 0 > gpurun_out/r02am_bench_isobea_2gpu_balance$b.json 2> gpurun_out/r02am_balance$b.err
tail -c 300 gpurun_out/r02am_balance$b.err
done
python - <<'P'
import json
for b in (10, 0):
    j = json.loads(open(f"gpurun_out/r02am_bench_isobea_2gpu_balance{b}.json").read().strip().splitlines()[-1])
    print("balance", b, "value %.4e ms/step %.0f parity %s" % (j["value"], j["ms_per_step"], j.get("parity", {}).get("ok")), j["config"].get("balanced_shards"))
P
