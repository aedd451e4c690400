#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "batched or patchify" -p no:cacheprovider > gpurun_out/test_new_ops.log 2>&1; echo "== newops exit $?"; tail -n 5 gpurun_out/test_new_ops.log
timeout 900 python -m pytest tests/test_sd3_gpu.py -q -m gpu -x -s -p no:cacheprovider > gpurun_out/test_sd3.log 2>&1; echo "== sd3 exit $?"; tail -n 25 gpurun_out/test_sd3.log
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01.csv python tools/profile_step.py > gpurun_out/ncu_launches.log 2>&1; echo "== ncu launches exit $?"; tail -n 3 gpurun_out/ncu_launches.log
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm -s 200 -c 3 -o gpurun_out/prof_igemm_r01 python tools/profile_step.py > gpurun_out/ncu_full.log 2>&1; echo "== ncu full exit $?"; tail -n 3 gpurun_out/ncu_full.log
