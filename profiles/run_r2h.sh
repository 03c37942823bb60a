#!/bin/bash
# round 2, call h: sub-warp-group step kernel: parity tests of all mappings, mapping timings; pinned-ReLU gradient tests
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py tests/test_train_gemm_gpu.py tests/test_rollout_gpu.py -m gpu -q -x > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log; tail -6 gpurun_out/r2h_pytest.log
timeout 600 python profiles/step_mappings.py > gpurun_out/r2h_step_mappings.jsonl 2> gpurun_out/r2h_step_mappings.err; cat gpurun_out/r2h_step_mappings.jsonl | cut -c1-400; tail -3 gpurun_out/r2h_step_mappings.err
