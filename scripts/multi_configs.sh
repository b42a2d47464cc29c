#!/bin/bash
# configs 3-5 at N GPUs (device-resident metric): bash scripts/multi_configs.sh N
N=$1
for c in 3 4 5; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$c bench.py --config $c --gpus $N --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/bench_r02_c${c}_n$N.json 2> gpurun_out/bench_r02_c${c}_n$N.err
  tail -1 gpurun_out/bench_r02_c${c}_n$N.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($c, $N, d['ms_per_step'], d['value'], d['config']['kernel_path'], d['roofline']['all_kernels_ms'])" || tail -5 gpurun_out/bench_r02_c${c}_n$N.err
done
