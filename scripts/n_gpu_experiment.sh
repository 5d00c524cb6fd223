#!/bin/bash
# usage: N=2 bash scripts/n_gpu_experiment.sh  -- overhead of the result gather / NCCL channel count at N GPUs
N=${N:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4))"; }
run 29541 "default"
EEM_BENCH_GATHER=0 run 29542 "no-gather(metrics only)"
NCCL_MAX_NCHANNELS=2 run 29543 "max 2 channels"
NCCL_MAX_NCHANNELS=4 run 29544 "max 4 channels"
NCCL_MAX_NCHANNELS=8 run 29545 "max 8 channels"
