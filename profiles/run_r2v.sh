#!/bin/bash
# round 2, call v: ncu --set full (with source-level sampling) of the staged tg_linear at 128 -> 128 and 128 -> 1
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 72 -c 1 -f -o gpurun_out/r2v_tg_linear_128x128_staged python profiles/tg_bench.py > gpurun_out/ncu_v1.log 2>&1; tail -2 gpurun_out/ncu_v1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 212 -c 1 -f -o gpurun_out/r2v_tg_linear_128x1_staged python profiles/tg_bench.py > gpurun_out/ncu_v2.log 2>&1; tail -2 gpurun_out/ncu_v2.log
