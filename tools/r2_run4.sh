out=gpurun_out/r2_brick1
mkdir -p $out
timeout 300 python tests/mgpu_check.py 12000 6 brick oracle > $out/brick_w1_a.txt 2>&1; tail -3 $out/brick_w1_a.txt
timeout 300 python tests/mgpu_check.py 99999 30 brick oracle > $out/brick_w1_b.txt 2>&1; tail -3 $out/brick_w1_b.txt
