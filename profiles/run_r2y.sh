#!/bin/bash
# round 2, call y: full GPU suite, smoke, full bench line after the logits loss / stacked heads / staged tg_linear
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2y_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest_gpu.log; tail -6 gpurun_out/r2y_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2y_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2y_smoke.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2y_bench.json 2> gpurun_out/r2y_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -3 gpurun_out/r2y_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2y_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k})"
