#!/bin/bash
# round 2, call 4d (gpurun --gpus 2): the joint two-team optimizer step (one graph, one all-reduce for both teams)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_rollout_gpu.py -m gpu -q -x -k "joint or overlapped or graph_captured" > gpurun_out/r4d_pytest_joint.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r4d_pytest_joint.log; tail -15 gpurun_out/r4d_pytest_joint.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/r4d_dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r4d_dist_train_2gpu.log; tail -12 gpurun_out/r4d_dist_train_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4d_bench_2gpu.json 2> gpurun_out/r4d_bench_2gpu.err; echo "bench 2gpu rc=$?" | tee -a gpurun_out/r4d_bench_2gpu.err; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r4d_bench_2gpu.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/r4d_bench_2gpu.json')); print(d['value'], d['e2e']['value']); t=d['train']; print({k:t[k] for k in t if 'ms' in k or 'us' in k or 'spread' in k})"
