#!/bin/bash
# round 2, call 3k: TMA tensor-store epilogue of tg_linear: tests (bit-equal to the STG epilogue), per-shape times both ways
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r3k_pytest_tg.log 2>&1; tail -4 gpurun_out/r3k_pytest_tg.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r3k_tg_bench_tma.log 2>&1; grep tg_linear gpurun_out/r3k_tg_bench_tma.log
TG_BENCH_TMA_OUT=0 timeout 300 python profiles/tg_bench.py > gpurun_out/r3k_tg_bench_stg.log 2>&1; grep tg_linear gpurun_out/r3k_tg_bench_stg.log
