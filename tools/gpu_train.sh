#!/bin/bash
# 1-GPU visit for the training path: GEMM + backward tests, then the train workload with the GELU dual-store variants.
mkdir -p gpurun_out
for f in tests/test_gemm_gpu.py tests/test_backward_gpu.py tests/test_fullsize_gpu.py; do
  n=$(basename $f .py)
  ( timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_$n.log)"
done
for d in 1 0 3 1 0 3; do
  echo "TP_TRAIN_DUAL=$d $(TP_TRAIN_DUAL=$d timeout 300 python bench.py --workload train --steps 20 2>&1 | tail -1 | cut -c1-400)" | tee -a gpurun_out/train_ab.log
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -c 160 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 3 > gpurun_out/ncu_launches_train.log 2>&1
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -5 ) | tee gpurun_out/memcheck.log
