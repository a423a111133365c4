out=gpurun_out/r2_v2d
mkdir -p $out
timeout 300 python tools/v2_check.py 99999 2048 > $out/v2_check.txt 2>&1; grep -E "screen|fast vs|force evaluation|vs oracle|\[fast\]" $out/v2_check.txt | head -12
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 600 python bench.py --steps 50 > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; tail -3 $out/bench_n1_f64.err; python -c "
import json; d=json.load(open('$out/bench_n1_f64.json')); print(d['value'], d['ms_per_step'], d['parity'], d['roofline']['kernel_ms'], d['cpu_baseline'])"
