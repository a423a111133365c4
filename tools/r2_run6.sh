out=gpurun_out/r2_bench_a
mkdir -p $out
timeout 600 python bench.py --steps 100 > $out/bench_n1.json 2> $out/bench_n1.err; tail -2 $out/bench_n1.err
python - <<PY
import json
d=json.load(open('$out/bench_n1.json'))
print('N=1', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['roofline']['frac_executed'], d['roofline']['force_evaluation'], d['clocks'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 100 > $out/bench_n2.json 2> $out/bench_n2.err; tail -2 $out/bench_n2.err
python - <<PY
import json
d=json.load(open('$out/bench_n2.json'))
print('N=2', d['value'], d['ms_per_step'], d['e2e']['value'], d['parity']['ok'], d['config']['engine'])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus 2 --steps 100 --engine replicated --no-cpu-baseline > $out/bench_n2_repl.json 2> $out/bench_n2_repl.err
python - <<PY
import json
d=json.load(open('$out/bench_n2_repl.json'))
print('N=2 replicated', d['value'], d['ms_per_step'])
PY
