out=gpurun_out/r2_brick2
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 24000 20 brick oracle > $out/brick_w2_a.txt 2>&1; tail -4 $out/brick_w2_a.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/mgpu_check.py 99999 50 brick oracle > $out/brick_w2_b.txt 2>&1; tail -4 $out/brick_w2_b.txt
