#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU). Usage: scripts/profile_round.sh <tag>
# 1) launch list + DRAM bytes of every kernel of the bench command (cold-cache, serialised: compare SHARES)
# 2) --set full of representative launches of the dominant kernels
set -u
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 --episodes 0"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_launches.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_tcgen05 -s 200 -c 6 -o $OUT/${TAG}_gemm_full $CMD > $OUT/${TAG}_gemm_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pw_stream -s 20 -c 4 -o $OUT/${TAG}_stream_full $CMD > $OUT/${TAG}_stream_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw2_kernel -s 40 -c 4 -o $OUT/${TAG}_dw_full $CMD > $OUT/${TAG}_dw_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw5s_kernel -s 20 -c 4 -o $OUT/${TAG}_dw5s_full $CMD > $OUT/${TAG}_dw5s_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k "regex:mbs_kernel|mbx_kernel" -s 4 -c 2 -o $OUT/${TAG}_mbx_full $CMD > $OUT/${TAG}_mbx_full.log 2>&1
ls -la $OUT | tail -8
