#!/bin/bash
# ncu evidence for profiles/: launch list of one SDXL forward + full captures of the top kernels
mkdir -p gpurun_out
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_step.py > gpurun_out/ncu_l.log 2>&1; echo "== launches exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm -s 260 -c 3 -o gpurun_out/prof_igemm_final python tools/profile_step.py > gpurun_out/ncu_f1.log 2>&1; echo "== igemm full exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_kernel -s 20 -c 2 -o gpurun_out/prof_attn_final python tools/profile_step.py > gpurun_out/ncu_f2.log 2>&1; echo "== attn full exit $?"
timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"layernorm_kernel|gn_apply|gn_stats" -s 30 -c 4 -o gpurun_out/prof_norm_final python tools/profile_step.py > gpurun_out/ncu_f3.log 2>&1; echo "== norm full exit $?"
ls -la gpurun_out/*.ncu-rep
