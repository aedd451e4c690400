#!/bin/bash
# ncu evidence for profiles/: launch list of one eager SDXL forward + --set full captures of the dominant kernels.
mkdir -p gpurun_out
ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r01_final.csv python tools/profile_step.py > gpurun_out/ncu_l.log 2>&1; echo "launch list exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:igemm_kernel -s 200 -c 4 -f -o gpurun_out/prof_igemm_final python tools/profile_step.py > gpurun_out/ncu_f1.log 2>&1; echo "igemm exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:attn_ -s 40 -c 4 -f -o gpurun_out/prof_attn_final python tools/profile_step.py > gpurun_out/ncu_f2.log 2>&1; echo "attn exit $?"
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:layernorm|gn_|conv3x3_cin4" -s 30 -c 6 -f -o gpurun_out/prof_norm_final python tools/profile_step.py > gpurun_out/ncu_f3.log 2>&1; echo "norm exit $?"
ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:igemm --csv --log-file gpurun_out/igemm_dram.csv python tools/profile_step.py > gpurun_out/ncu_t.log 2>&1; echo "traffic exit $?"
ls -la gpurun_out/*.ncu-rep
