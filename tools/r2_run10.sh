out=gpurun_out/r2_f
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "coincident or neighbor" > $out/pytest_some.log 2>&1; tail -4 $out/pytest_some.log
( time timeout 900 python bench.py > $out/bench_n1.json 2> $out/bench_n1.err ) 2>&1 | grep real; tail -3 $out/bench_n1.err
python - <<PY
import json
d=json.loads([l for l in open('$out/bench_n1.json') if l.startswith('{')][-1])
print('N=1', '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], d['parity']['ok'], d['roofline']['frac'], d['roofline']['force_evaluation']['frac'])
print('preprocess', d.get('preprocess'))
print('million', d.get('million'))
print('cpu', d.get('cpu_baseline'), d.get('extras_note'))
PY
