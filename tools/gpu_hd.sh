#!/bin/bash
# 1-GPU visit for the batched tiling kernel: HD tests (bit-exact fixtures), the hd_tile bench record, one ncu capture
mkdir -p gpurun_out
for f in tests/test_hd_gpu.py tests/test_fullsize_gpu.py; do
  n=$(basename $f .py)
  ( timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_$n.log)"
done
( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_hd.err | tail -1 ) > gpurun_out/bench_hd.json
python -c "
import json
d = json.load(open('gpurun_out/bench_hd.json'))
print('hd_tile', d['hd_tile']); print('train', d['train']['fwd_bwd_ms'], 'ms_per_step', d['ms_per_step'])
"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hd_tile_batch -c 1 -f -o gpurun_out/prof_hd_tile \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hd_tile.log 2>&1
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page details > gpurun_out/prof_hd_tile.details.txt 2>/dev/null
grep -E "Duration|DRAM Throughput|Issue Slots Busy|Registers Per|Achieved Occupancy" gpurun_out/prof_hd_tile.details.txt
