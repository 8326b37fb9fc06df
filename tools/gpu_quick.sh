mkdir -p gpurun_out
( timeout 700 python -m pytest tests/test_backward_gpu.py -q -m gpu 2>&1 | tail -15 ) | tee gpurun_out/pytest_gpu.log
( timeout 200 python tools/train_step_once.py 2>&1 | tail -4 ) | tee gpurun_out/train_step.log
