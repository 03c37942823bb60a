#!/bin/bash
# round 2, call 3p: raw chunks of tg_linear by tensor-map loads: tg tests, per-shape times, update time + launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py tests/test_rollout_gpu.py -m gpu -q -x > gpurun_out/r3p_pytest.log 2>&1; tail -3 gpurun_out/r3p_pytest.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r3p_tg_bench.log 2>&1; grep tg_linear gpurun_out/r3p_tg_bench.log
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r3p_ppo_update_torch_profile.txt 2>&1; head -1 gpurun_out/r3p_ppo_update_torch_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r3p_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
