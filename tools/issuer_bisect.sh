# stress of the two-issuer folds: repeated full bench steps + full-size parity tests
for i in 1 2 3 4 5 6 7 8; do
  timeout 120 python bench.py --steps 6 --warmup 2 --no-cpu-baseline --no-e2e > /tmp/o.txt 2> /tmp/e.txt
  rc=$?
  echo "rc=$rc $(cut -c60-110 /tmp/o.txt | tail -1)"
  if [ $rc -ne 0 ]; then tail -4 /tmp/e.txt | cut -c1-900; break; fi
done
for i in 1 2 3; do timeout 300 python -m pytest tests/test_gpu_fullsize.py -x -q -m gpu 2>&1 | tail -1; done
