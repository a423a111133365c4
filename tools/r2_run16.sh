out=gpurun_out/r2_j
mkdir -p $out
timeout 900 python -m pytest tests/test_gpu_md10k.py -m gpu -q -x -s > $out/pytest_md10k.log 2>&1; tail -12 $out/pytest_md10k.log
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -4
