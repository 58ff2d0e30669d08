#!/bin/bash
# ncu evidence for workload C3 on one B200 (run from the repo root on the GPU box, after the same bench command has
# exited 0 without ncu):  bash tools/profile_c3.sh <tag>   ->  gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.ncu-rep
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra"
timeout 300 $CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu_launch.log 2>&1
# one whole minibatch of tensor-core GEMM + update launches, steady state (skip the first 3 minibatches: 18 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'gemm_tc_kernel|k_update' --launch-skip 62 --launch-count 18 \
  -f -o gpurun_out/${TAG}_full $CMD > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out/${TAG}_* | head
