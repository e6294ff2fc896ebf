#!/bin/bash
set -u
OUT=gpurun_out
mkdir -p $OUT
N=${1:-8}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r02ag_bench_n$N.json 2> $OUT/r02ag_bench_n$N.err; echo "bench n$N rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02ag_bench_n$N.json'))
print(d['value'], d['n_gpus'], d['gather_ms'], d['e2e'], d['ms_per_step'])
print({k: {kk: vv for kk, vv in v.items() if kk in ('evals_s','ms','us_per_leapfrog','us_per_leapfrog_with_gather','n_gpus','chains_per_gpu')} for k, v in d['configs'].items()})
PY
tail -3 $OUT/r02ag_bench_n$N.err
timeout 300 python -m pytest tests/test_gpu_nccl.py -m gpu -q 2>&1 | tail -3
