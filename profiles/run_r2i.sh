#!/bin/bash
# round 2, call i: select-style sub-warp-group step kernel: parity tests, mapping timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_step_gpu.py -m gpu -q -x > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log; tail -6 gpurun_out/r2i_pytest.log
timeout 600 python profiles/step_mappings.py > gpurun_out/r2i_step_mappings.jsonl 2> gpurun_out/r2i_step_mappings.err; cat gpurun_out/r2i_step_mappings.jsonl | cut -c1-400; tail -3 gpurun_out/r2i_step_mappings.err
