#!/bin/bash
# round 2, call p: ncu --set full of tg_linear at the 128 -> 128 and 128 -> 1 shapes (196 608 rows, inputs from HBM)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 72 -c 2 -f -o gpurun_out/r2p_tg_linear_128x128 python profiles/tg_bench.py > gpurun_out/ncu_p1.log 2>&1; tail -2 gpurun_out/ncu_p1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 212 -c 1 -f -o gpurun_out/r2p_tg_linear_128x1 python profiles/tg_bench.py > gpurun_out/ncu_p2.log 2>&1; tail -2 gpurun_out/ncu_p2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_wgrad -s 72 -c 1 -f -o gpurun_out/r2p_tg_wgrad_128x128 python profiles/tg_bench.py > gpurun_out/ncu_p3.log 2>&1; tail -2 gpurun_out/ncu_p3.log
