timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_run28.log 2>&1; tail -2 gpurun_out/pytest_run28.log
timeout 300 python tools/v2_check.py 99999 4096 2>&1 | grep -v "generic\]" | grep "mixed (\|eval2"
