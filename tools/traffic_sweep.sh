#!/bin/bash
# DRAM traffic of the C3 GEMM classes under raster band heights (VBNN_TC_GM) x L2 hint policies (VBNN_TC_L2HINT):
# ncu with three metrics only (one pass), one steady-state minibatch.  -> gpurun_out/traffic_gm<G>_h<H>.csv
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
for GM in ${GMS:-4 8 16}; do for H in ${HINTS:-0 1}; do
  VBNN_TC_GM=$GM VBNN_TC_L2HINT=$H timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
    --clock-control none -k regex:gemm_tc_kernel --launch-skip 42 --launch-count 14 --csv --log-file gpurun_out/traffic_gm${GM}_h${H}.csv \
    $CMD > /dev/null 2>&1
  echo "gm=$GM hint=$H rc=$?"
done; done
