out=gpurun_out/r2_n
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tests/mgpu_check.py 99999 60 brick oracle > $out/check_w2.txt 2>&1; grep "^brick" $out/check_w2.txt | tail -2
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29542 tools/brick_profile.py 99999 40 > $out/profile_w2.txt 2>&1; grep -v "Warn\|warn\|\*\*\*\|OMP" $out/profile_w2.txt | head -8
timeout 300 python tools/brick_profile.py 99999 20 2>&1 | grep -v "Warn\|warn" | head -4
