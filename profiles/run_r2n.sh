#!/bin/bash
# round 2, call n: chunked minibatch step of the fused update (L2-resident activations): tests + timings per chunk size
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_rollout_gpu.py tests/test_config1.py -m gpu -q -x > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log; tail -5 gpurun_out/r2n_pytest.log
timeout 900 python profiles/update_chunks.py > gpurun_out/r2n_update_chunks.log 2>&1; cat gpurun_out/r2n_update_chunks.log | tail -8
