#!/bin/bash
# round 2, call 3u: staged tg_wgrad with the block height forced to 64 / 32 rows
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
TG_BENCH_WGRAD_ROWS=64 timeout 300 python profiles/tg_bench.py > gpurun_out/r3u_tg_bench_wr64.log 2>&1; grep tg_wgrad gpurun_out/r3u_tg_bench_wr64.log
TG_BENCH_WGRAD_ROWS=32 timeout 300 python profiles/tg_bench.py > gpurun_out/r3u_tg_bench_wr32.log 2>&1; grep tg_wgrad gpurun_out/r3u_tg_bench_wr32.log
