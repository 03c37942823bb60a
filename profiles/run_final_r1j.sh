# Last GPU pass of round 1 (r1j): all GPU tests after the renderer + cleanups, smoke, a rendered sample frame from a live
# batch, and compute-sanitizer memcheck over the step / host-stream / renderer kernels.
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q --deselect tests/test_render_gpu.py > gpurun_out/r1j_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest_gpu.log; tail -3 gpurun_out/r1j_pytest_gpu.log
timeout 200 python -m pytest tests/test_render_gpu.py -q > gpurun_out/r1j_pytest_render.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r1j_pytest_render.log; tail -5 gpurun_out/r1j_pytest_render.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1j_smoke.log 2>&1; tail -2 gpurun_out/r1j_smoke.log
timeout 100 python profiles/render_sample.py > gpurun_out/r1j_render_sample.log 2>&1; tail -2 gpurun_out/r1j_render_sample.log
timeout 100 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_step_gpu.py tests/test_render_gpu.py -q -x \
    -k "(golden and env_3v3) or many_host_errors or (many_host and 1000 and 7) or render_matches" > gpurun_out/r1j_memcheck.log 2>&1; tail -4 gpurun_out/r1j_memcheck.log
