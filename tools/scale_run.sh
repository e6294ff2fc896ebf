#!/bin/bash
# N-GPU weak-scaling bench lines (run under gpurun --gpus 8): bash tools/scale_run.sh <tag>
TAG=${1:-rXX}
for N in 4 8; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
    bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_${TAG}_n$N.json 2> gpurun_out/bench_${TAG}_n$N.err
  echo "N=$N rc=$?"; cut -c1-160 gpurun_out/bench_${TAG}_n$N.json
done
