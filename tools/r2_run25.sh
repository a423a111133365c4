out=gpurun_out/r2_run25
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v Warn | head -7
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_run25/bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['parity']['ok'], d['parity']['max_err_over_tol'])
r=d['roofline']; print(r['frac'], r['frac_executed'], r['kernel_ms'], r['force_evaluation'])
print('fp32', d['fp32']['value'], d['fp32']['speedup_over_f64'], d['fp32']['parity']['ok'])
print(d['million']['strong_1e6']['value'], d['million']['weak_125k_per_gpu']['value'], d['preprocess']['value'])
PY
