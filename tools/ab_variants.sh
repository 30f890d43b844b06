#!/bin/bash
# A/B the row-kernel variants built by tools/build_variant.py: tools/ab_variants.sh name1 name2 ...
for v in "$@"; do
  echo "=== variant $v"
  export LARVANET_B200_LIB=$PWD/larvanet_b200/csrc/build/variants/lib_$v.so
  timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "row or chain" 2>&1 | tail -1
  timeout 120 python tools/row_vs_tile.py 32 8,270,480 32,270,480 64,64,64 2>&1 | tail -3
  timeout 120 python tools/row_stats.py 32 270 480 2>&1 | sed -n '1,7p'
done
