// fa_inst.cuh -- bodies of the per-NG launchers; included only by fa_inst_g<NG>.cu.
#pragma once
#include "fa_launch.h"

namespace fa {

#define FA_FOR_NA(X) X(1) X(2) X(3) X(4) X(5)

// p.pdl: launch with programmatic stream serialization, so that consecutive steps overlap their launch latency and
// action fetch with the predecessor's tail (see pdl_wait() in fa_kernels.cuh)
template <typename K, typename R>
static cudaError_t launch_one(K kernel, const StepParams<R> &p, int grid, int block, cudaStream_t stream) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &at;
    cfg.numAttrs = p.pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, p);
}

template <int NG, typename R>
cudaError_t launch_step_g(int na, bool many, int mapping, const StepParams<R> &p, int grid, int block, cudaStream_t stream) {
    cudaError_t e = cudaSuccess;
    switch (na) {
#define FA_CASE(NA)                                                                                   \
    case NA:                                                                                          \
        if (mapping == FA_KMAP_GROUP) {                                                               \
            if (many) e = launch_one(fa_step_group_kernel<NG, NA, R, true>, p, grid, block, stream);  \
            else e = launch_one(fa_step_group_kernel<NG, NA, R, false>, p, grid, block, stream);      \
        } else if (mapping == FA_KMAP_AGENT) {                                                        \
            if (many) e = launch_one(fa_step_wide_kernel<NG, NA, R, true>, p, grid, block, stream);   \
            else e = launch_one(fa_step_wide_kernel<NG, NA, R, false>, p, grid, block, stream);       \
        } else {                                                                                      \
            if (many) e = launch_one(fa_step_kernel<NG, NA, R, true>, p, grid, block, stream);        \
            else e = launch_one(fa_step_kernel<NG, NA, R, false>, p, grid, block, stream);            \
        }                                                                                             \
        break;
        FA_FOR_NA(FA_CASE)
#undef FA_CASE
    default: return cudaErrorInvalidValue;
    }
    return e != cudaSuccess ? e : cudaGetLastError();
}

template <int NG, typename R>
cudaError_t launch_reset_g(int na, const StateView<R> &st, const uint8_t *mask, R *obs, int E, uint64_t seed,
                           uint64_t env_id0, int grid, int block, cudaStream_t stream) {
    switch (na) {
#define FA_CASE(NA)                                                                                     \
    case NA: fa_reset_kernel<NG, NA, R><<<grid, block, 0, stream>>>(st, mask, obs, E, seed, env_id0); break;
        FA_FOR_NA(FA_CASE)
#undef FA_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

template <int NG, typename R> cudaError_t step_attr_g(int na, bool many, int mapping, cudaFuncAttributes *out) {
    switch (na) {
#define FA_CASE(NA)                                                                                            \
    case NA:                                                                                                   \
        if (mapping == FA_KMAP_GROUP)                                                                          \
            return many ? cudaFuncGetAttributes(out, (const void *)fa_step_group_kernel<NG, NA, R, true>)      \
                        : cudaFuncGetAttributes(out, (const void *)fa_step_group_kernel<NG, NA, R, false>);    \
        if (mapping == FA_KMAP_AGENT)                                                                          \
            return many ? cudaFuncGetAttributes(out, (const void *)fa_step_wide_kernel<NG, NA, R, true>)       \
                        : cudaFuncGetAttributes(out, (const void *)fa_step_wide_kernel<NG, NA, R, false>);     \
        return many ? cudaFuncGetAttributes(out, (const void *)fa_step_kernel<NG, NA, R, true>)                \
                    : cudaFuncGetAttributes(out, (const void *)fa_step_kernel<NG, NA, R, false>);
        FA_FOR_NA(FA_CASE)
#undef FA_CASE
    default: return cudaErrorInvalidValue;
    }
}

#define FA_INSTANTIATE(NG)                                                                                        \
    template cudaError_t launch_step_g<NG, float>(int, bool, int, const StepParams<float> &, int, int, cudaStream_t);  \
    template cudaError_t launch_step_g<NG, double>(int, bool, int, const StepParams<double> &, int, int, cudaStream_t); \
    template cudaError_t launch_reset_g<NG, float>(int, const StateView<float> &, const uint8_t *, float *, int,  \
                                                   uint64_t, uint64_t, int, int, cudaStream_t);                   \
    template cudaError_t launch_reset_g<NG, double>(int, const StateView<double> &, const uint8_t *, double *,    \
                                                    int, uint64_t, uint64_t, int, int, cudaStream_t);             \
    template cudaError_t step_attr_g<NG, float>(int, bool, int, cudaFuncAttributes *);                                 \
    template cudaError_t step_attr_g<NG, double>(int, bool, int, cudaFuncAttributes *);

}  // namespace fa
