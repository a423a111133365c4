out=gpurun_out/r2_m
mkdir -p $out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tools/brick_profile.py 99999 40 > $out/profile_w8.txt 2>&1; grep -v "Warn\|warn\|\*\*\*\|OMP" $out/profile_w8.txt | head -22
