out=gpurun_out/r2_v2a
mkdir -p $out
timeout 300 python tools/v2_check.py 99999 4096 > $out/v2_check.txt 2>&1; tail -25 $out/v2_check.txt
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline --steps 50 > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; tail -c 300 $out/bench_n1_f64.json
