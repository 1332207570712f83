TAG=r02b; OUT=gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --profile-steps 0 --episodes 0"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw2_kernel -s 40 -c 4 -o $OUT/${TAG}_dw_full $CMD > $OUT/${TAG}_dw_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:dw5s_kernel -s 20 -c 4 -o $OUT/${TAG}_dw5s_full $CMD > $OUT/${TAG}_dw5s_full.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:mbx_kernel -s 4 -c 2 -o $OUT/${TAG}_mbx_full $CMD > $OUT/${TAG}_mbx_full.log 2>&1
ls -la $OUT/*_full.ncu-rep
