#!/bin/bash
# A/B the chain-kernel variants built by tools/build_variant.py: tools/ab_chain.sh name1 name2 ...
for v in "$@"; do
  echo "=== variant $v"
  export LARVANET_B200_LIB=$PWD/larvanet_b200/csrc/build/variants/lib_$v.so
  timeout 120 python -m pytest tests/test_gpu_kernels.py -m gpu -x -q -k "chain" 2>&1 | tail -1
  timeout 120 python tools/row_vs_tile.py 32 16,48,48 1,180,320 16,64,64 8,270,480 2>&1 | tail -4
done
