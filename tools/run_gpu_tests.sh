#!/bin/bash
# Runs the GPU test groups in separate processes so that one hung kernel cannot take the whole session down.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
rc=0
for grp in "linear" "conv" "sdpa" "norm or groupnorm or rmsnorm" "timestep or elementwise or ddim or rope"; do
  name=$(echo "$grp" | awk '{print $1}')
  timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "$grp" -x --no-header -p no:cacheprovider > gpurun_out/test_$name.log 2>&1
  code=$?
  echo "== $name exit $code"; tail -n 15 gpurun_out/test_$name.log
  [ $code -ne 0 ] && rc=1
done
exit $rc
