#!/bin/bash
# round 2, call s: 16-loader-warp variant of tg_linear / tg_wgrad: tests, then profiles/tg_bench.py
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
V=emergent-multiagent-strategies_b200/variants/libfa_lw16.so
FORTATTACK_B200_LIB=$PWD/$V timeout 600 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x > gpurun_out/r2r_pytest_lw16.log 2>&1; tail -2 gpurun_out/r2r_pytest_lw16.log
FORTATTACK_B200_LIB=$PWD/$V timeout 300 python profiles/tg_bench.py > gpurun_out/r2r_tg_bench_lw16.log 2>&1; cat gpurun_out/r2r_tg_bench_lw16.log
timeout 900 python -m pytest tests/test_rollout_gpu.py tests/test_policy_gpu.py tests/test_config1.py -m gpu -q > gpurun_out/r2s_pytest.log 2>&1; tail -5 gpurun_out/r2s_pytest.log
