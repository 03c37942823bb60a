#!/bin/bash
# round 2, call z: torch profile + ncu launch list of the update after the logits loss / stacked heads / staged tg_linear
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r2z_ppo_update_torch_profile.txt 2>&1; head -1 gpurun_out/r2z_ppo_update_torch_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2z_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
ls -la gpurun_out | tail -4
