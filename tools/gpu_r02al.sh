#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 600 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x 2>&1 | tail -6
N=2
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > $OUT/r02al_bench_n${N}.json 2> $OUT/r02al_bench_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02al_bench_n2.json'))
print(d['value'], d['n_gpus'], 'gather_ms', d['gather_ms'], d['ms_per_step'], d['parity']['pass'], d['e2e']['value'])
print({k: {kk: vv for kk, vv in v.items() if kk in ('evals_s','ms','us_per_leapfrog','us_per_leapfrog_with_gather','n_gpus','chains_per_gpu')} for k, v in d['configs'].items() if k in ('cfg3','cfg5')})
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/r02al_bench_n2.err | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 2>/dev/null | cut -c1-300
