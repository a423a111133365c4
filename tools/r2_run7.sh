out=gpurun_out/r2_brick_prof
mkdir -p $out
timeout 300 python tools/brick_profile.py 99999 50 > $out/w1.txt 2>&1; cat $out/w1.txt | grep -v Warn | head -30
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29515 tools/brick_profile.py 99999 50 > $out/w2.txt 2>&1; grep -v "Warn\|\*\*\*\|OMP" $out/w2.txt | head -30
