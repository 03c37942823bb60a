#!/bin/bash
# round 2, call 3v: final-state pass on 1 GPU: GPU suite, smoke, both bench arms, launch list of the update, memcheck of the
# tcgen05 / TMA training kernels on small shapes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r3v_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3v_pytest_gpu.log; tail -4 gpurun_out/r3v_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r3v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r3v_smoke.log
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"; tail -2 gpurun_out/r3v_bench.err
python -c "
import json; d=json.load(open('gpurun_out/r3v_bench.json')); print(d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['frac_of_copy_ceiling']); r=d['rollout']; print({k:r[k] for k in r if 'ms' in k or 'us' in k}); t=d['train']; print(t['iteration_ms'], t['update_ms'])"
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r3v_bench_ref.json 2> gpurun_out/r3v_bench_ref.err; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 600 --csv --log-file gpurun_out/r3v_launches_update.csv \
    python profiles/prof_policy.py 16384 --ppo > gpurun_out/ncu_list.log 2>&1
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_train_gemm_gpu.py -m gpu -q -x \
    -k "staged_form and (1000 or 129 or 31 or 999) or block_heights and 33 or adam" > gpurun_out/r3v_memcheck_tg.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r3v_memcheck_tg.log; tail -6 gpurun_out/r3v_memcheck_tg.log
