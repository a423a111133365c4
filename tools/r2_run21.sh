out=gpurun_out/r2_run21
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -x -q > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 300 python tools/v2_check.py 99999 4096 > $out/fast_path_check.txt 2>&1; tail -25 $out/fast_path_check.txt
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v Warn | head -8
