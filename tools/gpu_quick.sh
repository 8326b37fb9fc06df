mkdir -p gpurun_out
( timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -8 ) | tee gpurun_out/pytest_gpu.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tp_gemm2 -s 4 -c 1 -f -o gpurun_out/prof_mlp0 \
    python tools/prof_shape.py 9216 4096 1024 1 > gpurun_out/ncu_mlp0.log 2>&1
tail -2 gpurun_out/ncu_mlp0.log
( timeout 200 python tools/gemm_phase_profile.py 2>&1 | grep -E "timed|busy" ) | tee gpurun_out/phase_profile.log
( timeout 400 python bench.py --no-cpu-baseline --no-e2e 2>&1 | tail -1 | cut -c1-400 ) | tee gpurun_out/bench_quick.log
( timeout 200 python tools/train_step_once.py 2>&1 | tail -6 ) | tee gpurun_out/train_step.log
