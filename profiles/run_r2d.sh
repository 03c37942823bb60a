#!/bin/bash
# round 2, call d: full GPU suite, full bench line, torch profile + ncu of the update on the tcgen05 training kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q -s > gpurun_out/r2d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest_gpu.log; tail -4 gpurun_out/r2d_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2d_bench.json 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2d_bench.json; tail -3 gpurun_out/r2d_bench.err
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r2d_ppo_update_torch_profile.txt 2>&1; head -1 gpurun_out/r2d_ppo_update_torch_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1200 --csv --log-file gpurun_out/r2d_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_ -s 60 -c 8 -f -o gpurun_out/r2d_prof_tg \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_tg.log 2>&1; tail -2 gpurun_out/ncu_tg.log
ls -la gpurun_out | tail -8
