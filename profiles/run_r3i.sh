#!/bin/bash
# round 2, call 3i: front_end / heads / loss kernels: rollout + policy + config tests, then the update timing of bench.py
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3i_pytest_gpu.log; tail -5 gpurun_out/r3i_pytest_gpu.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3i_bench.json 2> gpurun_out/r3i_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r3i_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3i_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k})"
