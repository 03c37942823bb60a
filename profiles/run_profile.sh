#!/bin/bash
# Runs on the GPU box (gpurun): smoke, bench, ncu launch list and full captures of the step kernel.
set -x
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err
python bench.py --impl reference --steps 300 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
    python bench.py --steps 20 --warmup 3 --quick --e2e-steps 5 --reps 1 > gpurun_out/ncu_list.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fa_step_kernel -s 5 -c 3 -f -o gpurun_out/prof_step_4096 \
    python bench.py --steps 20 --warmup 3 --quick --no-graph --e2e-steps 5 --reps 1 > gpurun_out/ncu_full_4096.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fa_step_kernel -s 5 -c 2 -f -o gpurun_out/prof_step_1m \
    python bench.py --envs 1048576 --steps 12 --warmup 3 --quick --no-graph --e2e-steps 3 --reps 1 > gpurun_out/ncu_full_1m.log 2>&1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.csv
ls -la gpurun_out
