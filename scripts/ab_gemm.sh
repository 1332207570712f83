#!/bin/bash
# dev A/B on the GPU box: row-streaming GEMM (csrc/gemm_stream.cu) against the tcgen05 kernel on the layer shapes it covers
for s in 0 1 0 1; do
  echo "== STREAM=$s"
  STREAM=$s timeout 300 python scripts/pw_bench.py 1024 1 2>&1 | grep -E "b0|b1|b2"
done
