# Round-2 late captures: the row-wise training kernels and the persistent DDPM loop with the L2 prefetch off (final default).
set -x
F="--set full --import-source on --clock-control none --profile-from-start off"
timeout 250 ncu $F -k 'regex:linear_rows|layernorm_fwd|layernorm_bwd' -s 8 -c 5 -f -o gpurun_out/r2_train_rows python tools/profile_step.py train > /dev/null 2>&1
timeout 250 ncu $F -k regex:cd_loop -c 1 -f -o gpurun_out/r2_cd_loop_nopf python tools/profile_step.py planner > /dev/null 2>&1
ls -la gpurun_out/r2_train_rows.ncu-rep gpurun_out/r2_cd_loop_nopf.ncu-rep
