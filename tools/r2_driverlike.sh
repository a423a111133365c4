out=gpurun_out/r2_driverlike
mkdir -p $out
python -c "import __graft_entry__ as g; g.build(); g.smoke()" 2>&1 | tail -2
( time python bench.py --gpus 1 --steps 20 --warmup 5 > $out/b1.json 2> $out/b1.err ) 2>&1 | grep real; python -c "
import json; d=json.loads([l for l in open('$out/b1.json') if l.startswith('{')][-1]); print({k:d[k] for k in ('metric','value','unit','n_gpus','steps','warmup','ms_per_step','scaling','dtype','gpu_launches')}, d['clocks'], d['config']['parallelism'])"
( time python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > $out/r1.json 2> $out/r1.err ) 2>&1 | grep real; head -c 400 $out/r1.json; echo
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 2 --steps 20 --warmup 5 > $out/b2.json 2> $out/b2.err ) 2>&1 | grep real; python -c "
import json; d=json.loads([l for l in open('$out/b2.json') if l.startswith('{')][-1]); print({k:d[k] for k in ('value','n_gpus','ms_per_step','gpu_launches')}, d['parity']['ok'], d['preprocess']['value'], d['million']['strong_1e6']['value'])"
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $out/r2.json 2> $out/r2.err ) 2>&1 | grep real; python -c "
import json; d=json.loads([l for l in open('$out/r2.json') if l.startswith('{')][-1]); print(d['value'], d['cpu_baseline'])"
