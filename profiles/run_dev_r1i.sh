# Dev pass r1i on the GPU box: new host-stream step call, folded message rounds (tests + update profile), quick bench.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_step_gpu.py -x -q -k "many_host or step_host" > gpurun_out/r1i_pytest_host.log 2>&1; tail -3 gpurun_out/r1i_pytest_host.log
timeout 600 python -m pytest tests/test_rollout_gpu.py -x -q > gpurun_out/r1i_pytest_rollout.log 2>&1; tail -3 gpurun_out/r1i_pytest_rollout.log
timeout 300 python bench.py --quick > gpurun_out/r1i_bench_quick.json 2> gpurun_out/r1i_bench_quick.err; tail -c 1500 gpurun_out/r1i_bench_quick.json; tail -3 gpurun_out/r1i_bench_quick.err
timeout 300 python profiles/prof_policy.py 16384 --ppo --tf32 > gpurun_out/r1i_ppo_update_tf32_torch_profile.txt 2>&1; head -3 gpurun_out/r1i_ppo_update_tf32_torch_profile.txt
timeout 300 python profiles/prof_policy.py 16384 --ppo --tf32 --nofold > gpurun_out/r1i_ppo_update_tf32_nofold_torch_profile.txt 2>&1; head -3 gpurun_out/r1i_ppo_update_tf32_nofold_torch_profile.txt
timeout 300 python profiles/prof_policy.py 16384 --ppo > gpurun_out/r1i_ppo_update_fp32_torch_profile.txt 2>&1; head -3 gpurun_out/r1i_ppo_update_fp32_torch_profile.txt
