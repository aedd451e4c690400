#!/bin/bash
# round-2 ncu evidence for profiles/: launch list of one eager SDXL forward + --set full captures per kernel family.
mkdir -p gpurun_out
N="ncu --profile-from-start off --clock-control none"
$N --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r02_launches.csv python tools/profile_step.py > gpurun_out/ncu_l.log 2>&1; echo "launch list exit $?"
$N --set full --import-source on -k regex:igemm_kernel -s 200 -c 4 -f -o gpurun_out/r02_igemm python tools/profile_step.py > gpurun_out/ncu_f1.log 2>&1; echo "igemm exit $?"
$N --set full --import-source on -k regex:attn_ -s 40 -c 4 -f -o gpurun_out/r02_attn python tools/profile_step.py > gpurun_out/ncu_f2.log 2>&1; echo "attn exit $?"
$N --set full --import-source on -k "regex:layernorm|gn_|conv3x3_cin4" -s 30 -c 8 -f -o gpurun_out/r02_norm python tools/profile_step.py > gpurun_out/ncu_f3.log 2>&1; echo "norm exit $?"
$N --set full --import-source on -k "regex:attn_kernel|small_attention|softmax_rows" -c 3 -f -o gpurun_out/r02_misc python tools/profile_misc.py > gpurun_out/ncu_f4.log 2>&1; echo "misc exit $?"
$N --metrics dram__bytes_read.sum,dram__bytes_write.sum -k regex:igemm --csv --log-file gpurun_out/r02_igemm_dram.csv python tools/profile_step.py > gpurun_out/ncu_t.log 2>&1; echo "traffic exit $?"
ls -la gpurun_out/r02_*
