#!/bin/bash
# round 2, call t: ncu --set full of the 16-loader-warp variant at 128 -> 128 and 128 -> 1 (compare with profiles/r2p_*)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
export FORTATTACK_B200_LIB=$PWD/emergent-multiagent-strategies_b200/variants/libfa_lw16.so
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 72 -c 1 -f -o gpurun_out/r2t_tg_linear_128x128_lw16 python profiles/tg_bench.py > gpurun_out/ncu_t1.log 2>&1; tail -2 gpurun_out/ncu_t1.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tg_linear -s 212 -c 1 -f -o gpurun_out/r2t_tg_linear_128x1_lw16 python profiles/tg_bench.py > gpurun_out/ncu_t2.log 2>&1; tail -2 gpurun_out/ncu_t2.log
