out=gpurun_out/r2_run27
mkdir -p $out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cutoff_distance or fast_path" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29572 bench.py --gpus 2 > $out/bench_n2.json 2> $out/bench_n2.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_run27/bench_n2.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['parity']['ok'], d['preprocess'])
PY
timeout 300 python bench.py --no-cpu-baseline --steps 20 > $out/bench_n1.json 2> $out/bench_n1.err; python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r2_run27/bench_n1.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','n_gpus','ms_per_step')}, d['preprocess']['value'], d['preprocess']['ms_total'])
PY
