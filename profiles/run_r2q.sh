#!/bin/bash
# round 2, call q: full bench line with the parity sample (wall time recorded)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
t0=$(date +%s); timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2q_bench.json 2> gpurun_out/r2q_bench.err; echo "bench rc=$? wall $(( $(date +%s) - t0 )) s"
python -c "
import json; d=json.load(open('gpurun_out/r2q_bench.json')); print(d['value'], d['roofline']['frac'], d['roofline']['traffic']); print(json.dumps(d['cpu_baseline'].get('parity_sample'))); print(d['rollout']['ppo_update_ms'], d['rollout']['ppo_update_graph_ms'])"
t0=$(date +%s); timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2q_bench_ref.json 2> gpurun_out/r2q_bench_ref.err; echo "ref rc=$? wall $(( $(date +%s) - t0 )) s"
