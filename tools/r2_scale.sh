# usage (8-GPU box): bash tools/r2_scale.sh <outdir>
out=${1:-gpurun_out/r2_scale}
mkdir -p $out
run() { # world, port, extra args...
  w=$1; port=$2; shift 2
  if [ $w = 1 ]; then timeout 600 python "$@"; else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $w --master-addr 127.0.0.1 --master-port $port "$@"; fi
}
run 8 29521 tests/mgpu_check.py 24000 20 brick oracle > $out/check_w8.txt 2>&1; grep "^brick" $out/check_w8.txt | tail -2
for w in 8 4 2 1; do
  run $w $((29530 + w)) bench.py --gpus $w --steps 100 --no-cpu-baseline > $out/bench_n$w.json 2> $out/bench_n$w.err
  python - <<PY
import json
try:
    d=json.loads([l for l in open('$out/bench_n$w.json') if l.startswith('{')][-1])
    print('N=$w', '%.4g'%d['value'], '%.4f ms'%d['ms_per_step'], 'e2e %.4g'%d['e2e']['value'], 'parity', d['parity']['ok'], d['config']['engine']['bricks'], d['config']['engine']['owned_atoms_max_over_ranks'])
except Exception as e: print('N=$w failed', e)
PY
done
run 8 29541 tools/brick_profile.py 99999 20 > $out/profile_w8.txt 2>&1; grep -v "Warn\|warn\|\*\*\*\|OMP" $out/profile_w8.txt | head -20
