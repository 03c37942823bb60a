# Runs on the GPU box (gpurun): the round-end sequence the driver runs (GPU tests, smoke, both bench arms) plus the ncu
# launch list of the bench command and a full capture of the policy kernel.
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r1h_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1h_pytest_gpu.log; tail -3 gpurun_out/r1h_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1h_smoke.log 2>&1; tail -2 gpurun_out/r1h_smoke.log
timeout 300 python bench.py --impl reference --steps 300 --warmup 10 > gpurun_out/r1h_bench_ref.json 2> gpurun_out/r1h_bench_ref.err
timeout 900 python bench.py > gpurun_out/r1h_bench.json 2> gpurun_out/r1h_bench.err; tail -c 600 gpurun_out/r1h_bench.json; tail -3 gpurun_out/r1h_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1h_launches.csv \
    python bench.py --steps 20 --warmup 3 --quick --e2e-steps 5 --reps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mp_policy_kernel -s 2 -c 1 -f -o gpurun_out/r1h_prof_policy_16384 \
    python profiles/prof_policy.py 16384 > gpurun_out/ncu_policy.log 2>&1
ncu --set full --clock-control none -k regex:attn_ -c 2 -f -o gpurun_out/r1h_prof_attn \
    python -m pytest tests/test_rollout_gpu.py -q -k "fused_training_attention and 3-3-128" > gpurun_out/ncu_attn.log 2>&1
ls gpurun_out | tail -12
