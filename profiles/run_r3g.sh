#!/bin/bash
# round 2, call 3g: e2e (fa_step_many_host chunk pipeline) against the chunk size and the steps per call
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for mb in 12 16 24 32 64; do for ke in 256 1024; do
  FA_HOST_CHUNK_MB=$mb timeout 300 python bench.py --steps 20 --warmup 5 --quick --e2e-steps $ke > gpurun_out/r3g_b.json 2> gpurun_out/r3g_b.err
  python -c "
import json; d=json.load(open('gpurun_out/r3g_b.json')); e=d['e2e']; print('chunk_mb $mb steps/call $ke: e2e %.3e probe %.3e frac %.3f mapped %s' % (e['value'], e['copy_only_probe']['value'], e['frac_of_copy_ceiling'], e.get('mapped_single_launch', e.get('mapped'))))" | tee -a gpurun_out/r3g_e2e_chunks.log
done; done
