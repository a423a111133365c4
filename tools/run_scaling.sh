# usage (inside gpurun --gpus 8): bash tools/run_scaling.sh [list of N, default "1 2 4 8"]   -> gpurun_out/scaling/*.json
mkdir -p gpurun_out/scaling
export PANTEA_DIST_TIMEOUT_S=120
NS=${1:-"1 2 4 8"}
run() { # n atoms tag
  if [ "$1" = 1 ]; then timeout 300 python bench.py --gpus 1 --steps 50 --warmup 3 --atoms $2 --no-cpu-baseline > gpurun_out/scaling/$3.json 2> gpurun_out/scaling/$3.err
  else timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $1 --master-addr 127.0.0.1 --master-port $((29500 + $1)) bench.py --gpus $1 --steps 50 --warmup 3 --atoms $2 --no-cpu-baseline > gpurun_out/scaling/$3.json 2> gpurun_out/scaling/$3.err; fi
  tail -n 1 gpurun_out/scaling/$3.json | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])" || tail -3 gpurun_out/scaling/$3.err
}
for n in $NS; do run $n 100000 n${n}_100k; done
run 8 1000000 n8_1m
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/mgpu_check.py 48000 4 2>&1 | grep -E "world=|rror" | head -3
