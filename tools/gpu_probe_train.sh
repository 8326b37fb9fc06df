mkdir -p gpurun_out
( timeout 300 python tools/peer_store_probe.py 2>&1 | tail -2 ) | tee gpurun_out/peer_store_probe.log
( TP_SCHEDULE=4 timeout 300 python tools/peer_store_probe.py 2>&1 | tail -1 ) | tee -a gpurun_out/peer_store_probe.log
for f in tests/test_backward_gpu.py tests/test_fullsize_gpu.py tests/test_projector_gpu.py; do
  n=$(basename $f .py)
  ( timeout 900 python -m pytest $f -q -m gpu 2>&1 | tail -40 ) > gpurun_out/pytest_$n.log
  echo "$n: $(tail -1 gpurun_out/pytest_$n.log)"
done
rm -f gpurun_out/train_ab.log
for d in default 0 default 0; do
  if [ $d = default ]; then unset TP_TRAIN_DUAL; else export TP_TRAIN_DUAL=$d; fi
  echo "TP_TRAIN_DUAL=$d $(timeout 300 python bench.py --workload train --steps 20 2>&1 | tail -1 | cut -c1-500)" | tee -a gpurun_out/train_ab.log
done
unset TP_TRAIN_DUAL
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -s 150 -c 150 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train --steps 3 > gpurun_out/ncu_launches_train.log 2>&1
