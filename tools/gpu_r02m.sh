#!/bin/bash
# two GPUs: the NCCL test of the sharded path and the bench under torchrun
set -u
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi -L > $OUT/r02m_gpus.txt
timeout 900 python -m pytest tests/test_gpu_nccl.py -m gpu -q -x > $OUT/r02m_pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/r02m_pytest.log
tail -5 $OUT/r02m_pytest.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/r02m_bench_n2.json 2> $OUT/r02m_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02m_bench_n2.json'))
print(d['value'], d['n_gpus'], d['gather_ms'], d['e2e'], d['config']['collective'])
print({k: {kk: vv for kk, vv in v.items() if kk in ('evals_s','ms','us_per_leapfrog','us_per_leapfrog_with_gather','n_gpus','chains_per_gpu')} for k, v in d['configs'].items()})
PY
tail -5 $OUT/r02m_bench_n2.err
