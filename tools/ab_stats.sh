#!/bin/bash
# timing-only A/B (variants whose results are garbage): tools/ab_stats.sh name1 name2 ...
for v in "$@"; do
  echo "=== variant $v"
  export LARVANET_B200_LIB=$PWD/larvanet_b200/csrc/build/variants/lib_$v.so
  timeout 120 python tools/row_stats.py 32 270 480 2>&1 | sed -n "1,14p"
done
