#!/bin/bash
# ncu evidence for profiles/ (run under gpurun, ONE GPU). Usage: scripts/profile_round.sh <tag> [kernel regexes to capture in full ...]
# 1) launch list + DRAM bytes of every kernel of the bench command (cold-cache, serialised: compare SHARES)
# 2) --set full of representative launches of the dominant kernels, exported to CSV on the box (the .ncu-rep files of a whole round
#    exceed what gpurun copies back)
set -u
TAG=${1:-r01}
shift
OUT=gpurun_out
mkdir -p $OUT
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 --episodes 0"
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
    --csv --log-file $OUT/${TAG}_launches.csv $CMD > $OUT/${TAG}_launches.log 2>&1
full() {   # name, kernel regex, launches to skip, launches to capture
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s $3 -c $4 -o /tmp/${TAG}_$1 $CMD > $OUT/${TAG}_$1_full.log 2>&1
    ncu -i /tmp/${TAG}_$1.ncu-rep --page raw --csv > $OUT/${TAG}_$1_full_raw.csv 2>/dev/null
    rm -f /tmp/${TAG}_$1.ncu-rep
}
for k in "${@:-gemm stream dw dw5s mbx se}"; do
    case $k in
        gemm) full gemm pw_tcgen05 200 6 ;;
        stream) full stream pw_stream 20 4 ;;
        dw) full dw dw2_kernel 40 4 ;;
        dw5s) full dw5s dw5s_kernel 20 4 ;;
        mbx) full mbx "mbs_kernel|mbx_kernel" 4 2 ;;
        se) full se se_gate_kernel 48 16 ;;
    esac
done
ls -la $OUT | tail -8
