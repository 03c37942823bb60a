#!/bin/bash
# round 2, call 3c: tg_wgrad with 32-row blocks for 256-wide operands: tests, per-shape times with both heights
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r3c_pytest_tg.log 2>&1; tail -3 gpurun_out/r3c_pytest_tg.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r3c_tg_bench.log 2>&1; grep tg_wgrad gpurun_out/r3c_tg_bench.log
TG_BENCH_WGRAD_ROWS=64 timeout 300 python profiles/tg_bench.py > gpurun_out/r3c_tg_bench_wr64.log 2>&1; grep tg_wgrad gpurun_out/r3c_tg_bench_wr64.log
TG_BENCH_WGRAD_ROWS=32 timeout 300 python profiles/tg_bench.py > gpurun_out/r3c_tg_bench_wr32.log 2>&1; grep tg_wgrad gpurun_out/r3c_tg_bench_wr32.log
