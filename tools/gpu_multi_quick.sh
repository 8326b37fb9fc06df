#!/bin/bash
# Multi-GPU correctness visit:  gpurun --gpus N -- 'bash tools/gpu_multi_quick.sh N'   (dist tests + the configs[4] record with its bit-exactness check)
N=${1:-2}
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/pytest_dist_${N}gpu.log
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --workload hd5 --steps 50 --warmup 10 2>gpurun_out/bench_${N}gpu.err | tail -1 ) > gpurun_out/hd5_${N}gpu.json
python -c "
import json
d = json.load(open('gpurun_out/hd5_${N}gpu.json'))['hd5']
print({k: d.get(k) for k in ('fused_peer_store_ms', 'nccl_allgather_ms', 'rank_local_compute_ms', 'one_gpu_ms', 'strong_scaling_efficiency_fused', 'fused_bit_identical_to_one_gpu')})
" || tail -20 gpurun_out/bench_${N}gpu.err
