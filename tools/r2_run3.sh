out=gpurun_out/r2_v2e
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -8 $out/pytest_gpu.log
timeout 600 python bench.py --steps 50 --no-cpu-baseline --kernel-times > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; grep kernel-times $out/bench_n1_f64.err | head -12; python -c "
import json; d=json.load(open('$out/bench_n1_f64.json')); print(d['value'], d['ms_per_step'], d['parity']['ok'], d['roofline']['kernel_ms'])"
