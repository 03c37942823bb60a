set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tests/dist_train_gpu.py > gpurun_out/dist_train_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/dist_train_2gpu.log; tail -8 gpurun_out/dist_train_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2000 --warmup 10 > gpurun_out/bench_2gpu.json 2> gpurun_out/bench_2gpu.err; tail -c 1500 gpurun_out/bench_2gpu.json; tail -3 gpurun_out/bench_2gpu.err
timeout 300 python bench.py --impl reference --steps 300 --warmup 10 > gpurun_out/bench_ref.json 2>gpurun_out/bench_ref.err; tail -c 600 gpurun_out/bench_ref.json
