/*
 * fortattack_probe.h -- hardware probes and diagnostics of the tcgen05 building blocks.  TEST INFRASTRUCTURE:
 * built into libfortattack_probe.so (mp_probe_*) and, for mp_set_trace, into the diagnostic build
 * libfortattack_b200_trace.so (`make -C csrc trace`); none of it is part of the product library.
 */
#ifndef FORTATTACK_PROBE_B200_H
#define FORTATTACK_PROBE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Hardware probe of the tensor-core building block (one 128 x N x K tile); used by the tests. */
int mp_probe_gemm(const void *d_A, const void *d_Bp, float *d_out, int K, int N, uint32_t lbo, uint32_t sbo,
                  uint32_t idesc, uint32_t *d_err, void *stream);

/* Diagnostics (profiles/policy_phase_trace.py, profiles/umma_bulkcopy_microbench.py); not needed to run the policy.
 * mp_set_trace: device buffer of 96 uint64 that receives clock64() stamps of one row thread (entries 0..63), of the weight
 * producer (64..69) and of the MMA issuer (72..77) for one tile of CTA 0 (environment MP_TRACE_TILE selects which);
 * NULL switches tracing off.  mp_probe_timing: back-to-back tcgen05.mma issue rate for a 128 x N x 128 product and
 * cp.async.bulk latency / throughput with `depth` copies in flight; d_out int64 [3] = cycles of the three experiments. */
int mp_set_trace(unsigned long long *d_trace);
int mp_probe_timing(const void *d_Bp, int N, int reps, int bytes, int depth, long long *d_out, uint32_t *d_err, int grid,
                    void *stream);

#ifdef __cplusplus
}
#endif
#endif
