#!/bin/bash
# round 2, call 3r: fold_weights (rl_small_matmul): GPU suite, bench (update time), launch list of the update
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3r_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3r_pytest_gpu.log; tail -4 gpurun_out/r3r_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3r_bench.json 2> gpurun_out/r3r_bench.err; echo "bench rc=$?"; tail -2 gpurun_out/r3r_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3r_bench.json')); print(d['value'], d['e2e']['value']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); t=d['train']; print(t['iteration_ms'], t['update_ms'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r3r_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
