// mp_probe.cu -- a one-tile tcgen05 GEMM (128 x N x K, fp16 in, fp32 out) that exercises exactly the
// primitives and operand layout of the fused policy kernel (mp_umma.cuh).  tests/test_policy_gpu.py runs
// it against a float64 product; it is how the descriptor encoding is pinned on real hardware.
#include "mp_umma.cuh"

namespace mp {

// A: fp16 [128][K] row-major (global).  Bp: fp16 weights already in canonical layout, N x K
// (core matrix (n/8, k/8) at ((n/8) * (K/8) + k/8) * 128 bytes).  out: fp32 [128][N].
__global__ void __launch_bounds__(128) probe_gemm_kernel(const __half *A, const __half *Bp, float *out, int K, int N,
                                                         uint32_t lbo, uint32_t sbo, uint32_t idesc, uint32_t *err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_b, bar_mma;
    __shared__ uint32_t tmem_base_s;
    uint8_t *sA = smem;                          // 128 x K fp16
    uint8_t *sB = smem + 128 * K * 2;            // N x K fp16
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(&bar_b, 1);
        mbar_init(&bar_mma, 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (tid == 0) {
        mbar_expect_tx(&bar_b, (uint32_t)(N * K * 2));
        bulk_g2s(sB, Bp, (uint32_t)(N * K * 2), &bar_b);
    }
    // thread r stages row r of A in canonical layout (the way the policy kernel's epilogues write activations)
    const uint32_t a_lbo = 128u, a_sbo = (uint32_t)(K / 8) * 128u;
    for (int kc = 0; kc < K / 8; ++kc) {
        const uint4 v = *reinterpret_cast<const uint4 *>(A + (size_t)tid * K + kc * 8);
        *reinterpret_cast<uint4 *>(sA + canon_off(tid, kc * 8, a_lbo, a_sbo)) = v;
    }
    fence_async_smem();
    __syncthreads();

    if (tid == 0) {
        mbar_wait(&bar_b, 0, err, 1);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
        for (int k = 0; k < K / 16; ++k) {
            // one MMA consumes K=16 = two core matrices along K, which are a_lbo = 128 bytes apart
            const uint64_t ad = smem_desc(a0 + (uint32_t)k * 2u * a_lbo, lbo, sbo);
            const uint64_t bd = smem_desc(b0 + (uint32_t)k * 2u * a_lbo, lbo, sbo);
            umma_f16(tmem, ad, bd, idesc, k > 0);
        }
        umma_commit(&bar_mma);
    }
    mbar_wait(&bar_mma, 0, err, 2);
    tc_fence_after();
    for (int c = 0; c < N; c += 16) {
        uint32_t v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) out[(size_t)tid * N + c + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}


// Timing probe: (a) back-to-back tcgen05.mma throughput for a [128 x K=128] x [N x 128] product repeated
// `reps` times, (b) latency of one cp.async.bulk of `bytes`, (c) throughput with `depth` copies in flight.
// out[0] = cycles (a), out[1] = cycles (b, reps sequential copies), out[2] = cycles (c, reps copies, depth in flight)
__global__ void __launch_bounds__(128) probe_timing_kernel(const __half *Bp, int N, int reps, int bytes, int depth,
                                                           long long *out, uint32_t *err) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_mma, bar_ld[8];
    __shared__ uint32_t tmem_base_s;
    uint8_t *sA = smem;                    // 32 KB (contents irrelevant for timing)
    uint8_t *sB = smem + 32768;            // up to 64 KB
    uint8_t *sL = smem + 32768 + 65536;    // 8 x 16 KB copy targets
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        mbar_init(&bar_mma, 1);
        for (int i = 0; i < 8; ++i) mbar_init(&bar_ld[i], 1);
        mbar_fence_init();
    }
    if (warp == 0) tmem_alloc<256>(&tmem_base_s);
    for (int i = tid; i < (32768 + 65536) / 16; i += 128) reinterpret_cast<uint4 *>(smem)[i] = make_uint4(0, 0, 0, 0);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    if (tid == 0) {
        const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB), idesc = idesc_f16(128, N);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r)
            for (int k = 0; k < 8; ++k)
                umma_f16(tmem, smem_desc(a0 + k * 256u, 128, 2048), smem_desc(b0 + k * 256u, 128, 2048), idesc, k > 0);
        umma_commit(&bar_mma);
        mbar_wait(&bar_mma, 0, err, 1);
        out[0] = clock64() - t0;
        // (b) sequential copies
        uint32_t ph = 0;
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            mbar_expect_tx(&bar_ld[0], (uint32_t)bytes);
            bulk_g2s(sL, reinterpret_cast<const uint8_t *>(Bp) + (size_t)(r % 8) * bytes, (uint32_t)bytes, &bar_ld[0]);
            mbar_wait(&bar_ld[0], ph, err, 2);
            ph ^= 1u;
        }
        out[1] = clock64() - t0;
        // (c) `depth` copies in flight
        uint32_t phs[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        t0 = clock64();
        for (int r = 0; r < reps + depth; ++r) {
            const int s = r % depth;
            if (r >= depth) { mbar_wait(&bar_ld[s], phs[s], err, 3); phs[s] ^= 1u; }
            if (r < reps) {
                mbar_expect_tx(&bar_ld[s], (uint32_t)bytes);
                bulk_g2s(sL + s * 16384, reinterpret_cast<const uint8_t *>(Bp) + (size_t)(r % 8) * bytes, (uint32_t)bytes, &bar_ld[s]);
            }
        }
        out[2] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc<256>(tmem);
}

}  // namespace mp

extern "C" int mp_probe_gemm(const void *d_A, const void *d_Bp, float *d_out, int K, int N, uint32_t lbo, uint32_t sbo,
                             uint32_t idesc, uint32_t *d_err, void *stream) {
    if (K % 16 != 0 || N % 16 != 0 || N > 256 || K > 256 || N < 16) return -1;
    if (idesc == 0) idesc = mp::idesc_f16(128, N);
    const size_t smem = (size_t)(128 + N) * K * 2;
    cudaError_t e = cudaFuncSetAttribute(mp::probe_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    mp::probe_gemm_kernel<<<1, 128, smem, (cudaStream_t)stream>>>((const __half *)d_A, (const __half *)d_Bp, d_out, K, N, lbo,
                                                                 sbo, idesc, d_err);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}

extern "C" int mp_probe_timing(const void *d_Bp, int N, int reps, int bytes, int depth, long long *d_out, uint32_t *d_err,
                               int grid, void *stream) {
    if (N % 16 != 0 || N > 256 || N < 16 || bytes > 16384 || bytes % 16 != 0 || depth < 1 || depth > 8) return -1;
    const size_t smem = 32768 + 65536 + 8 * 16384;
    cudaError_t e = cudaFuncSetAttribute(mp::probe_timing_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return -2;
    mp::probe_timing_kernel<<<grid, 128, smem, (cudaStream_t)stream>>>((const __half *)d_Bp, N, reps, bytes, depth, d_out, d_err);
    return cudaGetLastError() == cudaSuccess ? 0 : -2;
}
