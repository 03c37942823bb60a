#!/bin/bash
# round 2, call 3o: final-state pass on 1 GPU: full GPU suite, smoke, both bench arms, launch list + torch profile of the update,
# ncu --set full of tg_linear (TMA epilogue) at 128 -> 128
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3o_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3o_pytest_gpu.log; tail -4 gpurun_out/r3o_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3o_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3o_smoke.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r3o_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3o_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); t=d['train']; print(t['iteration_ms'], t['update_ms'])"
t0=$(date +%s); timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3o_bench_ref.json 2> gpurun_out/r3o_bench_ref.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"; tail -c 400 gpurun_out/r3o_bench_ref.json
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r3o_ppo_update_torch_profile.txt 2>&1; head -1 gpurun_out/r3o_ppo_update_torch_profile.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 1000 --csv --log-file gpurun_out/r3o_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 72 -c 1 -f -o gpurun_out/r3o_tg_linear_128x128 python profiles/tg_bench.py > gpurun_out/ncu_o1.log 2>&1; tail -1 gpurun_out/ncu_o1.log
