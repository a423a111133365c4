out=gpurun_out/r2_run26
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q > $out/pytest_gpu_2gpu.log 2>&1; tail -3 $out/pytest_gpu_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29571 tests/mgpu_check.py 24000 20 brick oracle 2>&1 | grep -v "^\*\*\|OMP_NUM" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 > $out/bench_n2.json 2> $out/bench_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_run26/bench_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['parity']['ok'], d['fp32']['value'], d['million']['strong_1e6']['value'], d['preprocess']['value'])
PY
