# usage (inside gpurun, one GPU): bash tools/refresh_profiles.sh   -> gpurun_out/refresh/*
# Regenerates every artefact that profiles/ summarises: GPU test log, bench lines (FP64, FP32, reference arm),
# the ncu launch list of the bench command and one `--set full` capture of the three dominant kernels.
out=gpurun_out/refresh
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/pytest_gpu.log 2>&1; tail -1 $out/pytest_gpu.log
timeout 600 python bench.py > $out/bench_n1_f64.json 2> $out/bench_n1_f64.err; tail -c 400 $out/bench_n1_f64.json
timeout 300 python bench.py --dtype f32 --no-cpu-baseline > $out/bench_n1_f32.json 2> $out/bench_n1_f32.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $out/bench_reference_arm.json 2> $out/bench_reference_arm.err
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --kernel-times > /dev/null 2> $out/kernel_times.txt
timeout 300 python tools/bench_preprocess.py > $out/preprocess_batch.json 2> $out/preprocess_batch.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_bench_100k.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hdnnp_eval_kernel|pair_filter_kernel|neighbor_rows_kernel" \
    -s 6 -c 3 -o $out/prof_full -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $out/ncu_full.log 2>&1
# neighbour-scan variants of the same build (half-width cells / 5x5x5 stencil against the 3x3x3 scan)
PANTEA_CELL_STENCIL=1 timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline > $out/bench_n1_f64_coarse_cells.json 2> /dev/null
# halo-exchange path on one rank (its N > 1 numbers: torchrun ... bench.py --gpus N --halo on, and tests/mgpu_check.py ... halo)
timeout 200 python bench.py --steps 50 --warmup 3 --no-cpu-baseline --halo on > $out/bench_n1_halo.json 2> /dev/null
timeout 100 python tools/halo_probe.py > $out/halo_probe.txt 2>&1
ls -la $out
