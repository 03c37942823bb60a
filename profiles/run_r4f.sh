#!/bin/bash
# round 2, call 4f: final-state pass on 1 GPU with the two-stream team updates: GPU suite, smoke, both bench arms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
t0=$(date +%s); timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r4f_pytest_gpu.log 2>&1; echo "pytest rc=$? wall $(( $(date +%s) - t0 )) s" >> gpurun_out/r4f_pytest_gpu.log; tail -4 gpurun_out/r4f_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r4f_smoke.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s" | tee -a gpurun_out/r4f_bench.err; tail -2 gpurun_out/r4f_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r4f_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); t=d['train']; print(t['iteration_ms'], t['update_ms'], t['recompute_old_and_gae_ms'])"
t0=$(date +%s); timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r4f_bench_ref.json 2> gpurun_out/r4f_bench_ref.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"
