#!/bin/bash
# round 2, call o: tg_linear loader with double-buffered global loads: gemm tests, per-shape timings, update timing
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -q -x > gpurun_out/r2o_gemm.log 2>&1; tail -3 gpurun_out/r2o_gemm.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r2o_tg_bench.log 2>&1; grep tg_linear gpurun_out/r2o_tg_bench.log
timeout 300 python profiles/prof_policy.py 16384 --ppo 2>&1 | head -1
