#!/bin/bash
# round 2, call x: pipelined epilogue + redux amax; staged (bulk-copy) form of tg_linear: tests, then profiles/tg_bench.py in both forms
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r2x_pytest_tg.log 2>&1; tail -4 gpurun_out/r2x_pytest_tg.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r2x_tg_bench_staged.log 2>&1; grep tg_linear gpurun_out/r2x_tg_bench_staged.log
TG_BENCH_STAGED=0 timeout 300 python profiles/tg_bench.py > gpurun_out/r2x_tg_bench_regs.log 2>&1; grep tg_linear gpurun_out/r2x_tg_bench_regs.log
