#!/bin/bash
# round 2, call 3e: which of the two policy-kernel measures breaks the multi-tile case (MP_OPT bit 0 = balanced heads, bit 1 = pre-encode)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for o in 0 1 2 3; do
  MP_OPT=$o timeout 300 python -m pytest tests/test_policy_gpu.py -m gpu -q -x -k "blob_arithmetic" > gpurun_out/r3e_pytest_opt$o.log 2>&1; echo "MP_OPT=$o: $(tail -1 gpurun_out/r3e_pytest_opt$o.log)"
done
