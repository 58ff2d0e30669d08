#!/bin/bash
# compute-sanitizer passes over the small parity tests and the C++ replay host (the mbarrier / TMEM
# pipelines of gemm_tc.cu and the peer flags are where races would be).  Run on the GPU box from the repo root:
#   bash tools/sanitize.sh [outdir]          -> <outdir>/san_<tool>_<case>.log, summarised in profiles/rNN_sanitizer.md
# Each case is one pytest node id (or the replay binary) chosen to cover: tcgen05 single-CTA tiles, forced CTA-pair tiles
# (cta_group::2, cluster launch control, TMA multicast barriers), the multi-sample TMEM-resident dW, the elementwise kernels.
OUT=${1:-gpurun_out}
mkdir -p "$OUT"
CS=/usr/local/cuda/bin/compute-sanitizer
PY="python -m pytest -x -q -m gpu -p no:cacheprovider"
declare -A CASES=(
  [fused_bf16_lrt]="tests/test_gpu_mlp.py::test_fused_step_vs_oracle[local-False-bf16]"
  [fused_bf16_weight]="tests/test_gpu_mlp.py::test_fused_step_vs_oracle[weight-False-bf16]"
  [pair_tiles_lrt]="tests/test_gpu_mlp.py::test_pair_tile_epilogues_vs_oracle[local-1-1-0]"
  [pair_tiles_weight]="tests/test_gpu_mlp.py::test_pair_tile_epilogues_vs_oracle[weight-3-1-0]"
  [layer_golden]="tests/test_gpu_kernels.py::test_layer_api_against_golden"
)
run() {  # tool case command...
  local tool=$1 name=$2; shift 2
  local log="$OUT/san_${tool}_${name}.log"
  timeout 900 $CS --tool "$tool" --error-exitcode 99 --print-limit 20 "$@" > "$log" 2>&1
  local rc=$?
  echo "$tool $name rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|passed|failed' "$log" | tr '\n' ' ' | cut -c1-300)"
}
for tool in memcheck synccheck racecheck; do
  for name in "${!CASES[@]}"; do
    run $tool $name $PY "${CASES[$name]}"
  done
  run $tool replay ./tools/replay tests/golden/mlp_weight.replay.bin legacy
done
# initcheck: uninitialised device reads (padding columns of the operand buffers are the candidates)
run initcheck fused_bf16_lrt $PY "${CASES[fused_bf16_lrt]}"
