set -x
mkdir -p gpurun_out
export CUDA_LAUNCH_BLOCKING=1 FA_DBG_COLSUM=1
timeout 200 python profiles/dbg_dense.py update > gpurun_out/r1i_dbg_update.log 2>&1; grep -n "colsum\|Error" gpurun_out/r1i_dbg_update.log | head -40
