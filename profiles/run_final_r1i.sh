# Runs on the GPU box (gpurun): the round-end sequence the driver runs (GPU tests, smoke, both bench arms) plus the ncu
# launch list of the bench command, a full capture of the policy kernel, and full captures of the training kernels
# (folded-round attention forward/backward, ReLU-backward + bias gradient) at the config-3 minibatch shape.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1i_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1i_pytest_gpu.log; tail -3 gpurun_out/r1i_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1i_smoke.log 2>&1; tail -2 gpurun_out/r1i_smoke.log
timeout 300 python bench.py --impl reference --steps 300 --warmup 10 > gpurun_out/r1i_bench_ref.json 2> gpurun_out/r1i_bench_ref.err
timeout 900 python bench.py > gpurun_out/r1i_bench.json 2> gpurun_out/r1i_bench.err; tail -c 600 gpurun_out/r1i_bench.json; tail -3 gpurun_out/r1i_bench.err
timeout 300 python profiles/prof_policy.py 16384 --ppo --tf32 > gpurun_out/r1i_ppo_update_tf32_torch_profile.txt 2>&1; head -1 gpurun_out/r1i_ppo_update_tf32_torch_profile.txt
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r1i_ppo_update_fp32_torch_profile.txt 2>&1; head -1 gpurun_out/r1i_ppo_update_fp32_torch_profile.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1i_launches.csv \
    python bench.py --steps 20 --warmup 3 --quick --e2e-steps 5 --reps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mp_policy_kernel -s 2 -c 1 -f -o gpurun_out/r1i_prof_policy_16384 \
    python profiles/prof_policy.py 16384 > gpurun_out/ncu_policy.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k 'regex:attn_.*<4|relu_bwd_colsum' -s 2 -c 5 -f -o gpurun_out/r1i_prof_train_kernels \
    python profiles/prof_policy.py 16384 --ppo --tf32 > gpurun_out/ncu_train.log 2>&1; tail -2 gpurun_out/ncu_train.log
ls gpurun_out | tail -12
