#!/bin/bash
# round 2, call j: full GPU suite, full bench line (persistent headline), ncu of the persistent step kernel, smoke
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest_gpu.log; tail -6 gpurun_out/r2j_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2j_bench.json; tail -3 gpurun_out/r2j_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2j_smoke.log 2>&1; tail -2 gpurun_out/r2j_smoke.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fa_step_group_kernel -c 4 -f -o gpurun_out/r2j_prof_step \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_step.log 2>&1; tail -2 gpurun_out/ncu_step.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j_launches_bench_steps20.csv \
    python bench.py --steps 20 --warmup 5 --quick > gpurun_out/ncu_list2.log 2>&1
ls -la gpurun_out | tail -6
