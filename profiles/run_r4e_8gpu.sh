#!/bin/bash
# round 2, call 4e (gpurun --gpus 8): configs 4 and 5 at their full size: bench under torchrun exactly as the driver launches it
# (headline + e2e + train section with the gradient all-reduce + ensemble section), reference arm, NCCL checks of the training path
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "GPUs: $N"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r4e_bench_${N}gpu.json 2> gpurun_out/r4e_bench_${N}gpu.err; echo "bench rc=$?" | tee -a gpurun_out/r4e_bench_${N}gpu.err; tail -c 2500 gpurun_out/r4e_bench_${N}gpu.json; grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r4e_bench_${N}gpu.err | tail -4
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29562 bench.py --impl reference --gpus $N --steps 20 --warmup 5 > gpurun_out/r4e_bench_ref_${N}gpu.json 2> gpurun_out/r4e_bench_ref_${N}gpu.err; echo "ref rc=$?"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29563 tests/dist_train_gpu.py > gpurun_out/r4e_dist_train_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r4e_dist_train_${N}gpu.log; tail -2 gpurun_out/r4e_dist_train_${N}gpu.log
