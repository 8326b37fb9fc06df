mkdir -p gpurun_out
for f in tests/test_backward_gpu.py tests/test_fullsize_gpu.py; do
  n=$(basename $f .py)
  ( timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -30 ) > gpurun_out/pytest_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_$n.log)"
done
( timeout 500 compute-sanitizer --tool memcheck python tools/sanitize_small.py 2>&1 | tail -3 ) | tee gpurun_out/memcheck.log
