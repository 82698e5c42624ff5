#!/bin/bash
# Round-2 captures (one GPU): `ncu --set full` of the kernels north_star names + the new encoder kernel, and the launch list of
# one bench step.  Summaries: python tools/ncu_summary.py profiles/ncu_r02_kernels.csv profiles/r02_kernel_traffic.json gpurun_out/r02_*.ncu-rep
set -x
mkdir -p gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:gru_pass -s 6 -c 1 -f -o gpurun_out/r02_gru_pass python tools/bench_gru.py 32 2 > /dev/null 2>&1
ONLY_FUSED=1 $NCU -k regex:corr_pyramid -s 3 -c 1 -f -o gpurun_out/r02_corr_pyramid python tools/bench_corr.py 32 > /dev/null 2>&1
$NCU -k regex:"corr_lookup|lookup_conv" -s 2 -c 1 -f -o gpurun_out/r02_lookup python tools/profile_step.py > /dev/null 2>&1
$NCU -k regex:conv_rows -s 3 -c 1 -f -o gpurun_out/r02_conv_rows python tools/bench_rows.py 64 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_r02_step_b32_it8.csv python tools/profile_step.py > /dev/null 2>&1
ls -la gpurun_out/r02_*.ncu-rep gpurun_out/launches_r02_step_b32_it8.csv
