#!/bin/bash
# round 2, call 3l: tg_linear with the epilogue forms as separate instantiations: tg tests, per-shape times, full GPU suite, bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python profiles/tg_bench.py > gpurun_out/r3l_tg_bench_tma.log 2>&1; grep tg_linear gpurun_out/r3l_tg_bench_tma.log
TG_BENCH_TMA_OUT=0 timeout 300 python profiles/tg_bench.py > gpurun_out/r3l_tg_bench_stg.log 2>&1; grep tg_linear gpurun_out/r3l_tg_bench_stg.log | head -8
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3l_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3l_pytest_gpu.log; tail -4 gpurun_out/r3l_pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3l_bench.json 2> gpurun_out/r3l_bench.err; echo "bench rc=$?"; tail -3 gpurun_out/r3l_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3l_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k})"
