#!/bin/bash
# 1-GPU visit for the store-warp change (lane-indexed store jobs, whole-crop boxes): packed-output tests + the destination probe
mkdir -p gpurun_out
( timeout 300 python tools/peer_store_probe.py 2>&1 | tail -2 ) | tee gpurun_out/peer_store_probe_new.log
for f in tests/test_fullsize_gpu.py tests/test_hd_gpu.py tests/test_projector_gpu.py tests/test_gemm_gpu.py tests/test_backward_gpu.py; do
  n=$(basename $f .py)
  ( timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_$n.log)"
done
( timeout 300 python tools/peer_store_probe.py 2>&1 | tail -1 ) | tee -a gpurun_out/peer_store_probe_new.log
( timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | cut -c1-250 ) | tee gpurun_out/bench_quick.log
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4 ) | tee gpurun_out/memcheck.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hd_tile_batch -c 1 -f -o gpurun_out/prof_hd_tile \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hd_tile.log 2>&1
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page raw --csv > gpurun_out/prof_hd_tile.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page details > gpurun_out/prof_hd_tile.details.txt 2>/dev/null
