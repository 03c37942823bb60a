#!/bin/bash
# round 2, call a: full GPU suite (new: config-1 runs, trained-checkpoint parity, ratio test, validation) + quick bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/r2a_smi.txt 2>&1
nproc >> gpurun_out/r2a_smi.txt; numactl -H >> gpurun_out/r2a_smi.txt 2>&1; nvidia-smi topo -m >> gpurun_out/r2a_smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q -s --durations=15 > gpurun_out/r2a_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2a_pytest_gpu.log
tail -5 gpurun_out/r2a_pytest_gpu.log
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2a_bench_ref.json 2> gpurun_out/r2a_bench_ref.err
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
echo "bench rc=$?"
head -c 600 gpurun_out/r2a_bench.json
