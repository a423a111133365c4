out=gpurun_out/r2_run22
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -3 $out/pytest_gpu.log
timeout 300 python tools/v2_check.py 99999 4096 > $out/fast_path_check.txt 2>&1; grep -v "generic\]" $out/fast_path_check.txt | grep -v Warn | tail -22
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v Warn | head -6
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hdnnp_eval2_kernel|pair_filter2_kernel" -s 8 -c 2 -o $out/prof -f \
    python tools/v2_check.py 99999 0 > $out/ncu.log 2>&1
