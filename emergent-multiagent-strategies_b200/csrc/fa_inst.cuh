// fa_inst.cuh -- bodies of the per-NG launchers; included only by fa_inst_g<NG>.cu.
#pragma once
#include "fa_launch.h"

namespace fa {

#define FA_FOR_NA(X) X(1) X(2) X(3) X(4) X(5)

template <int NG, typename R>
cudaError_t launch_step_g(int na, bool many, bool wide, const StepParams<R> &p, int grid, int block, cudaStream_t stream) {
    switch (na) {
#define FA_CASE(NA)                                                                          \
    case NA:                                                                                 \
        if (wide) {                                                                          \
            if (many) fa_step_wide_kernel<NG, NA, R, true><<<grid, block, 0, stream>>>(p);   \
            else fa_step_wide_kernel<NG, NA, R, false><<<grid, block, 0, stream>>>(p);       \
        } else {                                                                             \
            if (many) fa_step_kernel<NG, NA, R, true><<<grid, block, 0, stream>>>(p);        \
            else fa_step_kernel<NG, NA, R, false><<<grid, block, 0, stream>>>(p);            \
        }                                                                                    \
        break;
        FA_FOR_NA(FA_CASE)
#undef FA_CASE
    default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
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

template <int NG, typename R> cudaError_t step_attr_g(int na, bool many, bool wide, cudaFuncAttributes *out) {
    switch (na) {
#define FA_CASE(NA)                                                                                            \
    case NA:                                                                                                   \
        if (wide)                                                                                              \
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
    template cudaError_t launch_step_g<NG, float>(int, bool, bool, const StepParams<float> &, int, int, cudaStream_t);  \
    template cudaError_t launch_step_g<NG, double>(int, bool, bool, const StepParams<double> &, int, int, cudaStream_t); \
    template cudaError_t launch_reset_g<NG, float>(int, const StateView<float> &, const uint8_t *, float *, int,  \
                                                   uint64_t, uint64_t, int, int, cudaStream_t);                   \
    template cudaError_t launch_reset_g<NG, double>(int, const StateView<double> &, const uint8_t *, double *,    \
                                                    int, uint64_t, uint64_t, int, int, cudaStream_t);             \
    template cudaError_t step_attr_g<NG, float>(int, bool, bool, cudaFuncAttributes *);                                 \
    template cudaError_t step_attr_g<NG, double>(int, bool, bool, cudaFuncAttributes *);

}  // namespace fa
