#!/bin/bash
# round 2, call 4a: the two teams' updates on two streams: bit-equality test, rollout test file, A/B timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x > gpurun_out/r4a_pytest_rollout.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4a_pytest_rollout.log; tail -5 gpurun_out/r4a_pytest_rollout.log
timeout 600 python profiles/update_overlap.py > gpurun_out/r4a_update_overlap.json 2> gpurun_out/r4a_update_overlap.err; echo "ab rc=$?"; cat gpurun_out/r4a_update_overlap.json; tail -3 gpurun_out/r4a_update_overlap.err
