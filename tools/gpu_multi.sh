#!/bin/bash
# Multi-GPU visit:  gpurun --gpus N -- 'bash tools/gpu_multi.sh N'
# 2-rank correctness test of the sharded / fused HD paths, then the driver's bench line at N GPUs (carries the configs[4] record).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_${N}gpu.txt 2>&1
( timeout 600 python -m pytest tests/test_dist_gpu.py -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/pytest_dist_${N}gpu.log
( timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/bench_${N}gpu.err | tail -1 ) > gpurun_out/bench_${N}gpu.json
python - <<PY
import json
try:
    d = json.load(open("gpurun_out/bench_${N}gpu.json"))
    print("N=${N} value", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e ms", d["e2e"] and round(d["e2e"]["ms_per_step"], 3), d["e2e"] and d["e2e"]["per_rank_h2d_gbs"])
    print(json.dumps(d["hd5"], indent=1))
except Exception as e:
    print("bench line missing:", e)
    print(open("gpurun_out/bench_${N}gpu.err").read()[-3000:])
PY
# NVLink evidence for the fused exchange: per-link data counters of GPU 0 before / after the configs[4] run (60 fused + 60 NCCL steps)
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_before_${N}gpu.txt 2>&1
( timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --workload hd5 --steps 50 --warmup 10 2>>gpurun_out/bench_${N}gpu.err | tail -1 ) > gpurun_out/hd5_${N}gpu.json
nvidia-smi nvlink -gt d -i 0 > gpurun_out/nvlink_after_${N}gpu.txt 2>&1
for sch in 3 1; do
  ( TP_SCHEDULE=$sch timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
      bench.py --gpus $N --workload hd5 --steps 50 --warmup 10 2>>gpurun_out/bench_${N}gpu.err | tail -1 ) > gpurun_out/hd5_${N}gpu_sched$sch.json
  python -c "
import json
d = json.load(open('gpurun_out/hd5_${N}gpu_sched$sch.json'))['hd5']
print('TP_SCHEDULE=$sch', {k: d.get(k) for k in ('fused_peer_store_ms', 'nccl_allgather_ms', 'rank_local_compute_ms', 'one_gpu_ms', 'strong_scaling_efficiency_fused', 'fused_bit_identical_to_one_gpu')})
"
done
cut -c1-1500 gpurun_out/hd5_${N}gpu.json
head -12 gpurun_out/nvlink_after_${N}gpu.txt
