#!/bin/bash
# dev A/B: time the expand + project shapes with each variant library
for v in "$@"; do
  echo "== variant $v"
  ORBIT_B200_LIB=/root/repo/variants/$v.so timeout 120 python scripts/pw_bench.py 128 1 2>&1 | grep -E "b1.0 exp|b1.1 proj|b4.1 proj|head"
done
