// fa_launch.h -- host-side entry points of the per-team-size kernel instantiations.
// Each fa_inst_g<NG>.cu translation unit instantiates the step / reset kernels for NG guards and
// 1..5 attackers in float and double, so the five units compile in parallel.
#pragma once
#include "fa_kernels.cuh"

// thread mappings of the step kernel (the values of include/fortattack.h FA_MAP_*)
enum { FA_KMAP_ENV = 1, FA_KMAP_AGENT = 2, FA_KMAP_GROUP = 3 };

namespace fa {

template <int NG, typename R>
cudaError_t launch_step_g(int na, bool many, int mapping, const StepParams<R> &p, int grid, int block, cudaStream_t stream);

template <int NG, typename R>
cudaError_t launch_reset_g(int na, const StateView<R> &st, const uint8_t *mask, R *obs, int E, uint64_t seed,
                           uint64_t env_id0, int grid, int block, cudaStream_t stream);

template <int NG, typename R> cudaError_t step_attr_g(int na, bool many, int mapping, cudaFuncAttributes *out);

}  // namespace fa
