#!/bin/bash
# Multi-GPU visit, tile-schedule comparison only:  gpurun --gpus N -- 'bash tools/gpu_multi_sched.sh N "0 4 6"'
N=${1:-8}
SCHEDS=${2:-"0 4 6"}
mkdir -p gpurun_out
for sch in $SCHEDS; do
  ( TP_SCHEDULE=$sch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --workload hd5 --steps 50 --warmup 10 2>>gpurun_out/bench_${N}gpu.err | tail -1 ) > gpurun_out/hd5_${N}gpu_sched$sch.json
  python -c "
import json
d = json.load(open('gpurun_out/hd5_${N}gpu_sched$sch.json'))['hd5']
print('TP_SCHEDULE=$sch', {k: d.get(k) for k in ('fused_peer_store_ms', 'nccl_allgather_ms', 'rank_local_compute_ms', 'one_gpu_ms', 'strong_scaling_efficiency_fused', 'fused_bit_identical_to_one_gpu')})
"
done
