#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -q -x > gpurun_out/r2e_gemm.log 2>&1; tail -5 gpurun_out/r2e_gemm.log
timeout 900 python -m pytest tests/test_rollout_gpu.py tests/test_config5_gpu.py -q -x -s > gpurun_out/r2e_rollout.log 2>&1; tail -5 gpurun_out/r2e_rollout.log
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r2e_ppo_update_torch_profile.txt 2>&1; head -1 gpurun_out/r2e_ppo_update_torch_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2e_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
