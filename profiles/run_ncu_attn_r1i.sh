# ncu --set full of the folded-round attention kernels at the config-3 minibatch shape (65536 envs x 3 agents x 128).
set -x
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:attn_ -s 2 -c 3 -f -o gpurun_out/r1i_prof_attn_mix \
    python profiles/prof_policy.py 16384 --ppo --tf32 > gpurun_out/ncu_attn_mix.log 2>&1; tail -3 gpurun_out/ncu_attn_mix.log
