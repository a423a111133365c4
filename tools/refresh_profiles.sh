# usage (inside gpurun, one GPU): bash tools/refresh_profiles.sh   -> gpurun_out/r2_refresh/*
# Regenerates every artefact that profiles/r2_* summarises: GPU test log, bench lines (default FP64 line with its fp32 /
# preprocess / million objects, reference arm), per-kernel times of the brick step, the ncu launch list of the bench
# command and one `--set full` capture of the dominant kernels (double and single evaluation, pair filter, rows).
out=gpurun_out/r2_refresh
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv > $out/gpu.txt; nproc >> $out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; tail -c 300 $out/bench_n1_f64.json
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
timeout 300 python tools/brick_profile.py 99999 20 > $out/kernel_times_brick_step.txt 2>&1
timeout 300 python tools/v2_check.py 99999 4096 > $out/fast_path_check.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_100k.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extra --no-parity > $out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on \
    -k regex:"hdnnp_eval2_kernel|hdnnp_eval2f_kernel|pair_filter2_kernel|neighbor_rows_kernel" -s 8 -c 8 -o $out/prof_full -f \
    python tools/v2_check.py 99999 0 > $out/ncu_full.log 2>&1
ls -la $out
