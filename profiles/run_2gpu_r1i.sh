# 2-GPU pass r1i (gpurun --gpus 2): NCCL validation of the training path after the folded-round update, update weak
# scaling 1 vs 2 GPUs on the same box, and both bench arms under torchrun as the driver launches them.
set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/r1i_dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/r1i_dist_train_2gpu.log; tail -4 gpurun_out/r1i_dist_train_2gpu.log
timeout 300 python profiles/update_scaling.py > gpurun_out/r1i_update_scaling.log 2>&1
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/update_scaling.py >> gpurun_out/r1i_update_scaling.log 2>&1; grep update_scaling gpurun_out/r1i_update_scaling.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2000 --warmup 10 > gpurun_out/r1i_bench_2gpu.json 2> gpurun_out/r1i_bench_2gpu.err; tail -c 1800 gpurun_out/r1i_bench_2gpu.json; tail -3 gpurun_out/r1i_bench_2gpu.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29514 bench.py --impl reference --gpus 2 --steps 300 --warmup 10 > gpurun_out/r1i_bench_ref_2gpu.json 2> gpurun_out/r1i_bench_ref_2gpu.err; tail -c 400 gpurun_out/r1i_bench_ref_2gpu.json
