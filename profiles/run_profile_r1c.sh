# Runs on the GPU box (gpurun): ncu launch list of the bench command, full captures of the policy and step kernels,
# torch.profiler table of one PPO update.
set -x
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1c_launches.csv \
    python bench.py --steps 20 --warmup 3 --quick --e2e-steps 5 --reps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mp_policy_kernel -s 2 -c 1 -f -o gpurun_out/r1c_prof_policy_16384 \
    python profiles/prof_policy.py 16384 > gpurun_out/ncu_policy.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fa_step -s 5 -c 1 -f -o gpurun_out/r1c_prof_step_4096 \
    python bench.py --steps 20 --warmup 3 --quick --no-graph --e2e-steps 5 --reps 1 > gpurun_out/ncu_step.log 2>&1
ncu --set full --clock-control none -k regex:gae_kernel -c 1 -f -o gpurun_out/r1c_prof_gae \
    python -m pytest tests/test_rollout_gpu.py -q -k fused_gae > gpurun_out/ncu_gae.log 2>&1
python profiles/prof_policy.py 16384 --ppo > gpurun_out/r1c_ppo_update_torch_profile.txt 2>&1
ls -la gpurun_out
