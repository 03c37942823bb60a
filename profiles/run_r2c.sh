#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/diag_train_grads.py > gpurun_out/r2c_diag.log 2>&1
tail -80 gpurun_out/r2c_diag.log
timeout 600 python -m pytest tests/test_train_gemm_gpu.py -q > gpurun_out/r2c_gemm.log 2>&1; tail -3 gpurun_out/r2c_gemm.log
