#!/bin/bash
# round 2, call 3t: tensor-map staged loaders of tg_wgrad: tests (bit-equal to the register form), per-shape times both ways
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r3t_pytest_tg.log 2>&1; tail -4 gpurun_out/r3t_pytest_tg.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r3t_tg_bench_staged.log 2>&1; grep tg_wgrad gpurun_out/r3t_tg_bench_staged.log
TG_BENCH_WGRAD_STAGED=0 timeout 300 python profiles/tg_bench.py > gpurun_out/r3t_tg_bench_regs.log 2>&1; grep tg_wgrad gpurun_out/r3t_tg_bench_regs.log
