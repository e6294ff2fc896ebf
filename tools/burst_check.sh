for dbg in 0 64; do echo "debug $dbg"; ELISA_B200_I8_DEBUG=$dbg python bench.py --steps 1 --warmup 1 --draws 113664 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['classes_ms_per_step']; n=d['config']['draws_per_gpu']/18944; print(d['config']['draws_per_gpu'], {k: round(v/n,3) for k,v in c.items()})"; done
python -m pytest tests/test_gpu_fold_engines.py -x -q -m gpu 2>&1 | tail -2
