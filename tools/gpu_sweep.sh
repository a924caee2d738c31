#!/bin/bash
# A/B timings of tuning builds (make lib SUFFIX=...) and environment knobs in one gpurun call.
# usage: gpu_sweep.sh "<suffix>[:ENV=VAL,...]" ...     ("-" = the product build)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
python tools/time_batch.py warmup > /dev/null 2>&1   # builds the cached tree file
for spec in "$@"; do
  suf="${spec%%:*}"; envs=""
  [[ "$spec" == *:* ]] && envs="${spec#*:}"
  [[ "$suf" == "-" ]] && suf=""
  ( export VR_LIB_SUFFIX="$suf"; IFS=','; for kv in $envs; do export "$kv"; done
    timeout 300 python tools/time_batch.py "${spec}" 2>&1 | grep -E "TIMING|Error|error" )
done | tee -a gpurun_out/sweep.log
