#!/bin/bash
# Runs the GPU test groups in separate processes so that one hung kernel cannot take the whole session down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
run() {
  name=$1; shift
  timeout 900 python -m pytest "$@" -q -m gpu -x --no-header -p no:cacheprovider -s > gpurun_out/test_$name.log 2>&1
  code=$?
  echo "== $name exit $code"; tail -n 25 gpurun_out/test_$name.log
  [ $code -ne 0 ] && rc=1
}
if [ "$1" != "unet" ]; then
  run linear tests/test_ops_gpu.py -k "linear"
  run conv tests/test_ops_gpu.py -k "conv"
  run sdpa tests/test_ops_gpu.py -k "sdpa"
  run norm tests/test_ops_gpu.py -k "norm"
  run misc tests/test_ops_gpu.py -k "timestep or elementwise or ddim or rope"
fi
run unet tests/test_unet_gpu.py
exit $rc
