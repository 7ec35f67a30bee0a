#!/bin/bash
# usage: r2_multi.sh N  -> torchrun bench on N GPUs (peer-memory exchange, then NCCL exchange)
N=$1; mkdir -p gpurun_out
run() { timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 "$@"; }
run > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"; cut -c1-2000 gpurun_out/r2_bench_n$N.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d.get('run'), d['roofline']['kernel'], round(d['roofline']['kernel_ms_avg'],4), 'e2e', round(d['e2e']['value'],1), 'launches', d['gpu_launches'])"
BBMPC_P2P=0 run > gpurun_out/r2_bench_n${N}_nccl.json 2> gpurun_out/r2_bench_n${N}_nccl.err; echo "rc=$?"; python -c "
import sys,json
for l in open('gpurun_out/r2_bench_n${N}_nccl.json'):
    if l.startswith('{'):
        d=json.loads(l); print(d['n_gpus'], round(d['value'],1), 'steps/s', round(d['ms_per_step'],3), 'ms', d.get('run'))"
tail -3 gpurun_out/r2_bench_n$N.err
