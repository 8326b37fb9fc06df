#!/bin/bash
# Round-2 GPU visit (1 GPU): the GPU suite as the driver runs it (one process, -x), smoke, A/B of the
# launch plans, the full bench line, phase profile, ncu launch list with DRAM bytes, sanitizers.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15 ) > gpurun_out/pytest_driver_style.log
tail -2 gpurun_out/pytest_driver_style.log
( timeout 900 python tools/tma_store_probe.py 2>&1 ) | tee gpurun_out/tma_store_probe.log
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
ab chain_nofuse TP_FUSE_ATTN=0
ab plain_7_launches TP_FUSE_ATTN=0 TP_CHAIN=0
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_line.json
cut -c1-400 gpurun_out/bench_line.json
( timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref.err | tail -1 ) > gpurun_out/bench_ref_line.json
cut -c1-300 gpurun_out/bench_ref_line.json
( timeout 300 python tools/gemm_phase_profile.py 2>&1 ) > gpurun_out/phase_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
    --clock-control none -c 40 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
TP_FUSE_ATTN=0 TP_CHAIN=0 timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
    --clock-control none -c 80 --csv --log-file gpurun_out/launches_plain.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches_plain.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s 150 -c 150 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 3 > gpurun_out/ncu_launches_train.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_gemm2 -s 5 -c 1 -f -o gpurun_out/prof_fused_forward \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
ncu -i gpurun_out/prof_fused_forward.ncu-rep --page raw --csv > gpurun_out/prof_fused_forward.raw.csv 2>/dev/null
timeout 600 ncu --set full --clock-control none --import-source on -k regex:hd_tile_batch -c 1 -f -o gpurun_out/prof_hd_tile \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_hd_tile.log 2>&1
ncu -i gpurun_out/prof_hd_tile.ncu-rep --page raw --csv > gpurun_out/prof_hd_tile.raw.csv 2>/dev/null
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | head -60 ) > gpurun_out/memcheck.log
tail -3 gpurun_out/memcheck.log
( timeout 500 compute-sanitizer --tool synccheck python tools/sanitize_small.py 2>&1 | tail -8 ) | tee gpurun_out/synccheck.log
( timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -30 ) > gpurun_out/racecheck.log
tail -3 gpurun_out/racecheck.log
ls -la gpurun_out | head -70
