set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_policy_gpu.py tests/test_rollout_gpu.py -x -q > gpurun_out/pytest_policy.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_policy.log
tail -15 gpurun_out/pytest_policy.log
timeout 900 python bench.py > gpurun_out/bench2.json 2> gpurun_out/bench2.err; tail -c 2500 gpurun_out/bench2.json; tail -5 gpurun_out/bench2.err
