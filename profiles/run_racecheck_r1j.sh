# compute-sanitizer racecheck over the kernels that exchange data through shared memory: the thread-per-agent step kernel
# (agent block staged in shared memory, three block barriers per step), the ReLU-backward column sums and the renderer.
set -x
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_step_gpu.py tests/test_render_gpu.py -q -x \
    -k "(golden and env_3v3 and agent) or (render_matches and 64)" > gpurun_out/r1j_racecheck_step.log 2>&1; tail -4 gpurun_out/r1j_racecheck_step.log
timeout 60 compute-sanitizer --tool racecheck --print-limit 5 python profiles/dbg_dense.py colsum > gpurun_out/r1j_racecheck_colsum.log 2>&1; tail -3 gpurun_out/r1j_racecheck_colsum.log
