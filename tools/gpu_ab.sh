#!/bin/bash
# Short 1-GPU visit: the GPU suite exactly as the driver runs it (one process, -x), smoke, A/B lines of experiments, both bench arms.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_driver_style.log
tail -2 gpurun_out/pytest_driver_style.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/smoke.log
ab() {  # name, env...
  name=$1; shift
  ( env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>&1 | tail -1 | python -c "
import sys, json
try:
    d = json.loads(sys.stdin.read()); print('$name', 'ms_per_step', round(d['ms_per_step'], 4), 'launches/step', d['gpu_launches'] / d['steps'], 'single_image_ms', d['configs0_single_image']['gpu_ms'])
except Exception as e:
    print('$name', 'FAILED', e)
" ) | tee -a gpurun_out/ab.log
}
rm -f gpurun_out/ab.log
ab fused_default X=1
ab prefetch_next TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/pf.so
ab fused_default_again X=1
ab prefetch_next_again TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/pf.so
ab subbatch2 TP_SCHEDULE=4
ab subbatch4 TP_SCHEDULE=6
for sch in 4 6; do
  ( TP_SCHEDULE=$sch timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_projector_gpu.py -q -m gpu 2>&1 | tail -3 ) | tee gpurun_out/pytest_sched$sch.log
done
( TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/pf.so timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_projector_gpu.py -q -m gpu 2>&1 | tail -3 ) | tee gpurun_out/pytest_pf.log
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_line.json
cut -c1-300 gpurun_out/bench_line.json
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 ) > gpurun_out/bench_ref_line.json
cut -c1-600 gpurun_out/bench_ref_line.json
