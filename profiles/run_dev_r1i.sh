# Dev pass r1i on the GPU box: all GPU tests, then the PPO update under torch.profiler (TF32 and fp32, folded rounds + dense functions).
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_pytest_gpu.log 2>&1; tail -5 gpurun_out/r1i_pytest_gpu.log
timeout 300 python profiles/prof_policy.py 16384 --ppo --tf32 > gpurun_out/r1i_ppo_update_tf32_torch_profile.txt 2>&1; head -1 gpurun_out/r1i_ppo_update_tf32_torch_profile.txt
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r1i_ppo_update_fp32_torch_profile.txt 2>&1; head -1 gpurun_out/r1i_ppo_update_fp32_torch_profile.txt
