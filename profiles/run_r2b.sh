#!/bin/bash
# round 2, call b: first run of the tcgen05 training GEMM kernels (tg_gemm.cu)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_train_gemm_gpu.py -q -s -x --deselect tests/test_train_gemm_gpu.py::test_dense_layers_match_library_path > gpurun_out/r2b_gemm.log 2>&1
echo "gemm rc=$?" >> gpurun_out/r2b_gemm.log
tail -25 gpurun_out/r2b_gemm.log
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -q -s > gpurun_out/r2b_gemm_all.log 2>&1
echo "gemm_all rc=$?" >> gpurun_out/r2b_gemm_all.log
tail -15 gpurun_out/r2b_gemm_all.log
timeout 900 python -m pytest tests/test_rollout_gpu.py -q -x > gpurun_out/r2b_rollout.log 2>&1
echo "rollout rc=$?" >> gpurun_out/r2b_rollout.log
tail -15 gpurun_out/r2b_rollout.log
