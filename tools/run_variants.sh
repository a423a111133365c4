# usage: bash tools/run_variants.sh "<defines variant 1>" "<defines variant 2>" ...   (runs GPU tests on the first)
py() { python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('RESULT', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; }
first=1
for v in "$@"; do
  PANTEA_DEFINES="$v" python -m pantea_b200.csrc.build --force > /dev/null 2>&1
  if [ $first = 1 ]; then timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -3; first=0; fi
  echo "variant: $v"; py
done
