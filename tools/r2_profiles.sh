# Round-2 profile captures (run under gpurun, 1 GPU): launch lists of one step of each workload, and one
# `ncu --set full` capture of the ghost kernel, the persistent DDPM loop and the training attention kernels.
set -x
mkdir -p gpurun_out
L="--metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv"
timeout 280 ncu $L --log-file gpurun_out/r2_act3d_launches.csv python tools/profile_step.py act3d > /dev/null 2>&1
timeout 280 ncu $L --log-file gpurun_out/r2_planner_launches.csv python tools/profile_step.py planner > /dev/null 2>&1
timeout 280 ncu $L --log-file gpurun_out/r2_train_launches.csv python tools/profile_step.py train > /dev/null 2>&1
F="--set full --import-source on --clock-control none --profile-from-start off"
timeout 280 ncu $F -k regex:xattn6 -s 1 -c 1 -f -o gpurun_out/r2_xattn6 python tools/profile_step.py act3d > /dev/null 2>&1
timeout 280 ncu $F -k regex:cd_loop -c 1 -f -o gpurun_out/r2_cd_loop python tools/profile_step.py planner > /dev/null 2>&1
timeout 280 ncu $F -k regex:attn_fwd_mma -s 2 -c 1 -f -o gpurun_out/r2_train_attn_fwd python tools/profile_step.py train > /dev/null 2>&1
timeout 280 ncu $F -k regex:attn_bwd_d.*_mma -c 2 -f -o gpurun_out/r2_train_attn_bwd python tools/profile_step.py train > /dev/null 2>&1
timeout 280 ncu $F -k regex:topk_cluster\|ctx_kv\|gather_tokens -c 6 -f -o gpurun_out/r2_geometry python tools/profile_step.py act3d > /dev/null 2>&1
ls -la gpurun_out/r2_*
