#!/bin/bash
# usage: N=2 bash scripts/n_gpu_experiment.sh  -- cost of the overlapped result gather at N GPUs and the effect of leaving SMs to NCCL
N=${N:-2}
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 3 --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$2', d['n_gpus'], round(d['value']), round(d['ms_per_step'],4))"; }
run 29541 "default"
EEM_BENCH_GATHER=0 run 29542 "no-gather(metrics only)"
EEM_TF32_MAX_SMS=140 run 29543 "GEMM on 140 SMs"
EEM_TF32_MAX_SMS=132 run 29544 "GEMM on 132 SMs"
EEM_TF32_MAX_SMS=116 run 29545 "GEMM on 116 SMs"
