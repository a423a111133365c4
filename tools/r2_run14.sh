out=gpurun_out/r2_i
mkdir -p $out
timeout 1200 python -m pytest tests -m gpu -q -x > $out/pytest_gpu.log 2>&1; tail -5 $out/pytest_gpu.log
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -8
timeout 300 python tools/v2_check.py 99999 2048 2>&1 | grep -E "mixed|fast vs|vs oracle|\[mixed32\] (pantea::hdnnp|void pantea::pair)"
