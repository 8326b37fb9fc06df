#!/bin/bash
# One GPU visit: tests, smoke, bench (ours + reference arm), config sweep, launch list, full ncu capture of the dominant
# kernel, compute-sanitizer memcheck.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -25 ) > gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/smoke.log
( timeout 600 python bench.py 2>&1 | tail -3 ) | tee gpurun_out/bench.log
( timeout 300 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 ) | tee gpurun_out/bench_ref.log
( timeout 300 python tools/config_sweep.py 2>&1 | tail -3 ) > gpurun_out/config_sweep.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tp_gemm2 -s 4 -c 1 -f -o gpurun_out/prof_kv0 \
    python tools/prof_shape.py 36864 2048 4096 1 > gpurun_out/ncu_kv0.log 2>&1
tail -2 gpurun_out/ncu_kv0.log
( timeout 400 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -4 ) | tee gpurun_out/memcheck.log
ls -la gpurun_out | head -40
