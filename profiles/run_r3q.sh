#!/bin/bash
# round 2, call 3q: final-state pass on 1 GPU (after the tensor-map raw loads and the shared-memory tile accesses): GPU suite, smoke, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3q_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3q_pytest_gpu.log; tail -4 gpurun_out/r3q_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3q_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3q_smoke.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r3q_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3q_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); t=d['train']; print(t['iteration_ms'], t['update_ms'])"
timeout 300 python profiles/tg_bench.py > gpurun_out/r3q_tg_bench.log 2>&1; grep "128, \"N\": 128\|tg_wgrad" gpurun_out/r3q_tg_bench.log | head -6
