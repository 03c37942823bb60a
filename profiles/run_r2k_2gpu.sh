#!/bin/bash
# round 2, call k (gpurun --gpus 2): NCCL checks of the training path, both bench arms under torchrun exactly as the driver
# launches them (exit codes recorded: the graph-replayed multi-rank optimizer step must tear down cleanly), tg kernel microbench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/r2k_dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r2k_dist_train_2gpu.log; tail -3 gpurun_out/r2k_dist_train_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_2gpu.json 2> gpurun_out/r2k_bench_2gpu.err; echo "bench 2gpu rc=$?" | tee -a gpurun_out/r2k_bench_2gpu.err; tail -c 1500 gpurun_out/r2k_bench_2gpu.json; tail -5 gpurun_out/r2k_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2k_bench_ref_2gpu.json 2> gpurun_out/r2k_bench_ref_2gpu.err; echo "ref rc=$?"; tail -c 300 gpurun_out/r2k_bench_ref_2gpu.json
timeout 300 python profiles/tg_bench.py > gpurun_out/r2k_tg_bench.log 2>&1; cat gpurun_out/r2k_tg_bench.log
