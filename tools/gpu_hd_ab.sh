#!/bin/bash
# A/B of the tiling kernel's rows-per-thread / unroll variants (build_ab/*.so), each with the bit-exactness tests
mkdir -p gpurun_out; rm -f gpurun_out/hd_ab.log
for v in default u4 u8 r16u4 r4u4 default; do
  if [ $v = default ]; then unset TOKENPACKER_B200_LIB_OVERRIDE; else export TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/$v.so; fi
  t=$(timeout 300 python -m pytest tests/test_hd_gpu.py -q -m gpu 2>&1 | tail -1)
  r=$(timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())['hd_tile']; print(round(d['ms'], 4), round(d['ms_public_call'], 4))")
  echo "$v: hd_tile ms (launch, public call) = $r | tests: $t" | tee -a gpurun_out/hd_ab.log
done
