out=gpurun_out/r2_g
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -4 $out/pytest_gpu.log
( time timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err ) 2>&1 | grep real; tail -3 $out/bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open('$out/bench_n1.json') if l.startswith('{')][-1])
print('N=1', '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['parity']['ok'], d['roofline']['frac'], d['roofline']['force_evaluation']['frac'])
print('fp32', d.get('fp32'))
PY
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -6
