set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 2000 --warmup 10 > gpurun_out/r1e_bench_${N}gpu.json 2> gpurun_out/r1e_bench_${N}gpu.err; tail -c 400 gpurun_out/r1e_bench_${N}gpu.json; tail -2 gpurun_out/r1e_bench_${N}gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29542 bench.py --impl reference --gpus $N --steps 300 --warmup 10 > gpurun_out/r1e_bench_ref_${N}gpu.json 2> gpurun_out/r1e_bench_ref_${N}gpu.err; tail -c 300 gpurun_out/r1e_bench_ref_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29543 tests/dist_train_gpu.py > gpurun_out/r1e_dist_train_${N}gpu.log 2>&1; tail -2 gpurun_out/r1e_dist_train_${N}gpu.log
