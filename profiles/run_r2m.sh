#!/bin/bash
# round 2, call m: scenario-callback tests (generic particle world), renderer tests, quick bench with the NUMA binding report
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_mw_gpu.py tests/test_render_gpu.py -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m_pytest.log; tail -15 gpurun_out/r2m_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 --quick > gpurun_out/r2m_bench_quick.json 2> gpurun_out/r2m_bench_quick.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/r2m_bench_quick.json')); print(d['value'], d['e2e']['value'], d['e2e']['host_binding'], d['e2e']['copy_only_probe']['value'])"
