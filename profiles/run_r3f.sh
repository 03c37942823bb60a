#!/bin/bash
# round 2, call 3f: mp_policy_kernel time with the two tile-overlap measures switched (MP_OPT bit 0 = balanced heads, bit 1 = pre-encode)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for o in 0 1 2 3 0 3; do MP_OPT=$o timeout 300 python profiles/policy_time.py 2>&1 | tee -a gpurun_out/r3f_policy_time.log; done
