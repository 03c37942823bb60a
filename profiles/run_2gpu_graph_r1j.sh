# 2-GPU pass r1j: the multi-rank optimizer step (with its NCCL all-reduces) replayed from one CUDA graph.
set -x
mkdir -p gpurun_out
timeout 100 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 profiles/update_scaling.py --graph > gpurun_out/r1j_update_scaling_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r1j_update_scaling_graph.log; grep "update_scaling\|rc=\|Error\|error" gpurun_out/r1j_update_scaling_graph.log | tail -5
timeout 60 python profiles/update_scaling.py --graph >> gpurun_out/r1j_update_scaling_graph.log 2>&1; grep "update_scaling" gpurun_out/r1j_update_scaling_graph.log | tail -1
