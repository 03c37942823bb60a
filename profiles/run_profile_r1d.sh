# Runs on the GPU box (gpurun): full GPU test suite, smoke, ncu launch list of the bench command, full capture of the v3 policy kernel
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r1d_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1d_pytest_gpu.log; tail -4 gpurun_out/r1d_pytest_gpu.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r1d_launches.csv \
    python bench.py --steps 20 --warmup 3 --quick --e2e-steps 5 --reps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:mp_policy_kernel -s 2 -c 1 -f -o gpurun_out/r1d_prof_policy_16384 \
    python profiles/prof_policy.py 16384 > gpurun_out/ncu_policy.log 2>&1
ls -la gpurun_out | tail -5
