out=gpurun_out/r2_run23
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -2 $out/pytest_gpu.log
timeout 300 python tools/v2_check.py 99999 4096 > $out/fast_path_check.txt 2>&1; grep -v "generic\]" $out/fast_path_check.txt | grep -v Warn | grep "entries\|fast vs\|eval2\|filter2"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"pair_filter2_kernel" -s 4 -c 1 -o $out/prof -f \
    python tools/v2_check.py 99999 0 > $out/ncu.log 2>&1
