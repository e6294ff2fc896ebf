#!/bin/bash
set -u
OUT=gpurun_out; mkdir -p $OUT
N=${1:-2}
for g in peer nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 --gather $g > $OUT/r02ai_bench_n${N}_$g.json 2> $OUT/r02ai_bench_n${N}_$g.err; echo "bench n$N $g rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/r02ai_bench_n${N}_$g.json'))
print(d['value'], d['n_gpus'], 'gather_ms', d['gather_ms'], d['ms_per_step'], d['parity']['pass'])
print(d['config']['collective'][:60])
print({k: {kk: vv for kk, vv in v.items() if kk in ('evals_s','ms','us_per_leapfrog','us_per_leapfrog_with_gather','n_gpus','chains_per_gpu')} for k, v in d['configs'].items() if k in ('cfg3','cfg5')})
PY
grep -v "OMP_NUM_THREADS\|\*\*\*\*" $OUT/r02ai_bench_n${N}_$g.err | tail -4
done
