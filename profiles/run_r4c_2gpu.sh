#!/bin/bash
# round 2, call 4c (gpurun --gpus 2): the multi-rank training path after the permutation / scratch-scope changes (NCCL checks),
# and the bench under torchrun exactly as the driver launches it
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/r4c_dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r4c_dist_train_2gpu.log; tail -3 gpurun_out/r4c_dist_train_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r4c_bench_2gpu.json 2> gpurun_out/r4c_bench_2gpu.err; echo "bench 2gpu rc=$?" | tee -a gpurun_out/r4c_bench_2gpu.err; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r4c_bench_2gpu.err | tail -4
python -c "
import json; d=json.load(open('gpurun_out/r4c_bench_2gpu.json')); print(d['value'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling']); print(json.dumps(d['train'])[:1500])"
