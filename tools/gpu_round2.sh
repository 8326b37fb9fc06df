#!/bin/bash
# Round-2 GPU visit (1 GPU): tests, smoke, bench, phase profile, ncu (launch list with DRAM bytes + full sets of the five GEMM
# launches of one step), sanitizers.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 ) > gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
( timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 ) | tee gpurun_out/smoke.log
( timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/bench.err | tail -1 ) > gpurun_out/bench_line.json
cut -c1-1500 gpurun_out/bench_line.json
( timeout 300 python tools/gemm_phase_profile.py 2>&1 ) > gpurun_out/phase_profile.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.max \
    --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tp_gemm2 -s 21 -c 5 -f -o gpurun_out/prof_step_gemms \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ncu -i gpurun_out/prof_step_gemms.ncu-rep --page raw --csv > gpurun_out/prof_step_gemms.raw.csv 2>/dev/null
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -5 ) | tee gpurun_out/memcheck.log
( timeout 500 compute-sanitizer --tool synccheck python tools/sanitize_small.py 2>&1 | tail -5 ) | tee gpurun_out/synccheck.log
( timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_small.py 2>&1 | tail -12 ) | tee gpurun_out/racecheck.log
# ---- A/B: the development build (store warps, chained persistent GEMM launches, front work, split-K wgrads) on the same box
if [ -f build_ab/dev.so ]; then
  export TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/dev.so
  ( timeout 1200 python -m pytest tests -q -m gpu -x 2>&1 | tail -40 ) > gpurun_out/pytest_dev.log
  tail -5 gpurun_out/pytest_dev.log
  ( timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/bench_dev.err | tail -1 ) > gpurun_out/bench_dev_line.json
  cut -c1-700 gpurun_out/bench_dev_line.json
  timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
      --clock-control none -c 60 --csv --log-file gpurun_out/launches_dev.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launches_dev.log 2>&1
  ( TP_CHAIN=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null | tail -1 | cut -c1-400 ) > gpurun_out/bench_dev_nochain.json
  cat gpurun_out/bench_dev_nochain.json
  unset TOKENPACKER_B200_LIB_OVERRIDE
fi
ls -la gpurun_out | head -70
