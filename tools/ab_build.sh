#!/bin/bash
# Build a variant of libtokenpacker_b200.so for A/B experiments:   tools/ab_build.sh NAME -DFLAG [-DFLAG2 ...]
#   -> build_ab/NAME.so ; run anything against it with  TOKENPACKER_B200_LIB_OVERRIDE=$PWD/build_ab/NAME.so python ...
# Known flags: -DTP_B_PREFETCH (weight tiles before griddepcontrol.wait), -DTP_PAIR_STAGES=5, -DTP_EPI_SUB_PAIRS=4|16,
#              -DTP_ROLE_LAYOUT=0, -DTP_GEMM_PROFILE (cycle counters, tools/gemm_phase_profile.py)
set -e
name=$1; shift
root=$(cd "$(dirname "$0")/.." && pwd)
mkdir -p "$root/build_ab"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-fvisibility=hidden \
    --expt-relaxed-constexpr -shared -cudart static "$@" -o "$root/build_ab/$name.so" "$root/tokenpacker_b200/csrc/tp_api.cu"
echo "built $root/build_ab/$name.so"
