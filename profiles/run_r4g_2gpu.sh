#!/bin/bash
# round 2, call 4g (gpurun --gpus 2): NCCL training checks with BatchedTrainer's defaults (graph replay on, joint two-team step)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/r4g_dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r4g_dist_train_2gpu.log; tail -4 gpurun_out/r4g_dist_train_2gpu.log
timeout 300 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x -k "two_gpu" > gpurun_out/r4g_pytest_two_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4g_pytest_two_gpu.log; tail -3 gpurun_out/r4g_pytest_two_gpu.log
