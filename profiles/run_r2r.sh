#!/bin/bash
# round 2, call r: tg_linear / tg_wgrad with 16 loader warps per CTA (variant library) against the committed 8-warp kernels:
# the tg tests on both libraries, then profiles/tg_bench.py on both
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=emergent-multiagent-strategies_b200/variants/libfa_lw16.so
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r2r_pytest_lw8.log 2>&1; tail -2 gpurun_out/r2r_pytest_lw8.log
FORTATTACK_B200_LIB=$PWD/$V timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r2r_pytest_lw16.log 2>&1; tail -2 gpurun_out/r2r_pytest_lw16.log
timeout 300 python profiles/tg_bench.py > gpurun_out/r2r_tg_bench_lw8.log 2>&1; cat gpurun_out/r2r_tg_bench_lw8.log
FORTATTACK_B200_LIB=$PWD/$V timeout 300 python profiles/tg_bench.py > gpurun_out/r2r_tg_bench_lw16.log 2>&1; cat gpurun_out/r2r_tg_bench_lw16.log
timeout 600 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x > gpurun_out/r2r_pytest_rollout.log 2>&1; tail -3 gpurun_out/r2r_pytest_rollout.log
