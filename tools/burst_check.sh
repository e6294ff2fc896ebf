for dbg in 0 32 96 8; do echo "debug $dbg"; ELISA_B200_I8_DEBUG=$dbg python bench.py --steps 1 --warmup 1 --draws 113664 --no-cpu-baseline --no-e2e | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); c=d['roofline']['classes_ms_per_step']; n=d['config']['draws_per_gpu']/18944; print(d['config']['draws_per_gpu'], {k: round(v/n,3) for k,v in c.items()})"; done
