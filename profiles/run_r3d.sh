#!/bin/bash
# round 2, call 3d: policy kernel with balanced heads + next-tile opponent encoder under the heads' MMAs: tests, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_rollout_gpu.py tests/test_config5_gpu.py -m gpu -q -x > gpurun_out/r3d_pytest.log 2>&1; tail -4 gpurun_out/r3d_pytest.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3d_bench.json 2> gpurun_out/r3d_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3d_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3d_bench.json')); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); print(json.dumps(r['policy_kernel'])[:700]); print(d['rollout_5v5'].get('us_per_rollout_step'), d['rollout_ensemble'].get('us_per_rollout_step'))"
