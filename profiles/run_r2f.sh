#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/diag_train_grads.py 24 > gpurun_out/r2f_diag24.log 2>&1
grep -v "^      " gpurun_out/r2f_diag24.log | tail -20
grep "encoder.0.weight\|update.0.weight\|value_head.0.weight" gpurun_out/r2f_diag24.log
