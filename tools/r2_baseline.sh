# round-2 baseline of the round-1 kernels on this pool: GPU tests, bench, launch list, ncu --set full of the three kernels
out=gpurun_out/r2_base
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $out/gpu.txt
nproc >> $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; tail -c 600 $out/bench_n1_f64.json
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --kernel-times > /dev/null 2> $out/kernel_times.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hdnnp_eval_kernel|pair_filter_kernel|neighbor_rows_kernel" \
    -s 6 -c 3 -o $out/prof_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
ls -la $out
