// tg_gemm.cu -- the dense products of the PPO update (MPNN.evaluate_actions forward / backward, mpnn.py:117-205,
// rlcore/algo/ppo.py:146-192) and its optimizer step as hand-written sm_100a kernels.  C ABI: include/fortattack_train.h.
//
//   tg_linear   out = act(x B^T + bias)     rows ~ 2e5, K, N <= 256: HBM-bound (512 B in + 512 B out per row and 128
//               columns); fp32 operands are split into fp16 hi/lo terms on the fly, three tcgen05.mma per product
//   tg_wgrad    dW = x^T y  (reduction over the rows): bf16 three-term split, six MMAs per product, operands consumed
//               straight from their row-major tiles through MN-major shared-memory descriptors
//   tg_adam_step  gradient-norm clip + Adam for all parameter tensors in two launches
//
// tg_linear_kernel<STAGED, TMA_OUT>, persistent, one CTA per SM, 544 threads (576 in the staged form):
//   warps 0-7   epilogue: thread = accumulator row (tensor-memory lane; warps w and w + 4 share a lane window and split the
//               16-column blocks).  TMA_OUT: scales, bias, ReLU in registers, the row's 16 values into the warp's swizzled
//               32 x 16 tile, ONE cp.async.bulk.tensor store per block (an added term arrives by tensor-map loads of the same
//               blocks).  Otherwise: 32 x 16 blocks transposed through shared memory and stored with STG.128.
//   warps 8-15  loaders: a 128-row x 128-column group of x, per-row amax -> power-of-two scale, hi = fp16(x s),
//               lo = fp16(x s - hi), 16-byte stores into the canonical no-swizzle K-major layout (8-row core matrices
//               contiguous: SBO = 128; k-chunk stride LBO = 2048 + 16: the pad makes the stores of a warp whose lanes run along k
//               bank-conflict free).  STAGED (K = 64 / 128): the rows come out of a shared-memory ring that the producer warp
//               fills with one tensor-map load per 32-row chunk; otherwise coalesced float4 global loads.
//   warp 16     MMA issuer (converged warp, one elected lane): per 64-column stage 4 k-steps x {hi.hi, hi.lo, lo.hi}
//   warp 17     (staged form) producer of the raw ring
//   ring of NST operand stages (full/empty mbarriers), two accumulator buffers in tensor memory (acc_full/acc_empty), weights
//   (fp16 hi/lo planes, packed by tg_pack_weight) resident in shared memory for the whole launch.
// K = 256 is handled as two 128-column groups with their own row scales and their own accumulators (summed in the
// epilogue), so the per-row scale never has to wait for more than 128 columns.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include "../../include/fortattack_train.h"
#include "mp_umma.cuh"

int fa_internal_fail(int code, const char *fmt, ...);

namespace tg {
using namespace mp;

constexpr int ROWS = 128, KC = 64;
constexpr uint32_t A_SBO = 128, A_LBO = ROWS * 16 + 16;           // bytes
constexpr uint32_t HALF = (KC / 8) * A_LBO, STAGE = 2 * HALF;     // 16512, 33024
#ifndef TG_LOAD_WARPS
#define TG_LOAD_WARPS 8                                           // 8 or 16 loader warps per CTA (measured: profiles/r2r_*)
#endif
constexpr int EPI_WARPS = 8, LOAD_WARPS = TG_LOAD_WARPS, LOAD_WARP0 = EPI_WARPS, MMA_WARP = EPI_WARPS + LOAD_WARPS;
static_assert(LOAD_WARPS == 8 || LOAD_WARPS == 16, "8 or 16 loader warps");
constexpr int RPW = ROWS / LOAD_WARPS, NBATCH = RPW / 8;          // tile rows per loader warp; batches of 2 passes x 4 rows
constexpr int THREADS = (MMA_WARP + 1) * 32, LOAD_THREADS = LOAD_WARPS * 32, THREADS_STAGED = THREADS + 32;
constexpr int MAX_NST = 6, SCALE_SLOTS = 2;
// staged form (K = 64 or 128): a producer warp streams raw fp32 rows of x into a ring of RAW_ROWS-row chunks with bulk copies
// (TMA engine: the requests do not pass through registers or L1TEX); the loader warps become converters that read the chunks
// from shared memory.  Chunk = one pass of the converters: RAW_ROWS = LOAD_WARPS x 4 rows.
constexpr int RAW_ROWS = LOAD_WARPS * 4, NRAW = 4, PROD_WARP = MMA_WARP + 1;
constexpr int TB_STRIDE = 20;                                     // floats per row of a warp's 32 x 16 transposition block
constexpr int TB_BYTES = EPI_WARPS * 32 * TB_STRIDE * 4, RS_BYTES = SCALE_SLOTS * 2 * ROWS * 4, BIAS_BYTES = 2 * 256 * 4;
constexpr int FIXED_BYTES = TB_BYTES + RS_BYTES + BIAS_BYTES;
constexpr int SMEM_LIMIT = 232448 - 1024;                         // 227 KB minus the static barriers

struct LinParams {
    const float *x;
    const uint8_t *packed;
    const float *bias;
    float *out;
    const float *res;                                             // acc: out = act(..) + res (row stride ldr); may alias out
    uint32_t *status;
    long long rows;
    int ldx, ldo, ldr, K, Kp, N, Np, relu, acc, n_tiles, nst, groups, gw, vec_ok, wbytes;
    int tma_out;                                                  // epilogue stores through the TMA engine (tensor map `omap`)
    int nraw, tb_bytes;                                           // chunks of the raw ring (staged form); bytes of the epilogue's tile area
};

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) { return __half22float2(*reinterpret_cast<const __half2 *>(&u)); }

// 8 consecutive fp32 (already scaled) -> 16 bytes of fp16 hi and 16 bytes of fp16 lo
__device__ __forceinline__ void split8_f16(const float4 &a, const float4 &b, uint4 &hi, uint4 &lo) {
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        h[j] = pack_h2(v[2 * j], v[2 * j + 1]);
        const float2 f = unpack_h2(h[j]);
        l[j] = pack_h2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
    }
    hi = make_uint4(h[0], h[1], h[2], h[3]);
    lo = make_uint4(l[0], l[1], l[2], l[3]);
}

__device__ __forceinline__ float amax4(const float4 &a) { return fmaxf(fmaxf(fabsf(a.x), fabsf(a.y)), fmaxf(fabsf(a.z), fabsf(a.w))); }
__device__ __forceinline__ void scale4(float4 &a, float s) { a.x *= s; a.y *= s; a.z *= s; a.w *= s; }

// power-of-two scale s with amax * s in [2^14, 2^15) (s = 1 for an all-zero row); returns s, *inv = 1 / s
__device__ __forceinline__ float pow2_scale(float amax, float *inv) {
    int e = (int)(__float_as_uint(amax) >> 23) - 127;
    e = amax > 0.0f ? max(-100, min(100, e)) : 14;
    *inv = __uint_as_float((uint32_t)(127 + e - 14) << 23);
    return __uint_as_float((uint32_t)(127 + 14 - e) << 23);
}

// ---- epilogue stores through the TMA engine -------------------------------------------------------------------------------
// ncu on the STG epilogue: one warp-wide STG.E.128 (512 bytes) costs ~34 wavefronts of the LSU data pipe, the output stores
// were 53 % of that pipe's work and the pipe the busiest unit of the kernel (70 %).  A 32-row x 16-column block leaves instead as
// ONE cp.async.bulk.tensor store from a dense shared-memory tile (rows of 64 bytes, SWIZZLE_64B: the 16-byte chunk j of row r
// lives at chunk j ^ ((r >> 1) & 3), which also makes the row-per-thread STS.128 conflict-free); rows past the matrix are
// clipped by the tensor map.
__device__ __forceinline__ void tma_store_2d(const CUtensorMap *map, uint32_t smem_addr, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(reinterpret_cast<uint64_t>(map)),
                 "r"(c0), "r"(c1), "r"(smem_addr)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap *map, uint32_t smem_addr, int c0, int c1, uint64_t *bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_addr),
                 "l"(reinterpret_cast<uint64_t>(map)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <bool STAGED, bool TMA_OUT>
__global__ void __launch_bounds__(STAGED ? THREADS_STAGED : THREADS, 1) tg_linear_kernel(const LinParams p, const __grid_constant__ CUtensorMap omap, const __grid_constant__ CUtensorMap rmap,
                 const __grid_constant__ CUtensorMap xmap) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NST], bar_empty[MAX_NST], bar_acc_full[2], bar_acc_empty[2], bar_w;
    __shared__ __align__(8) uint64_t bar_raw_full[NRAW], bar_raw_empty[NRAW], bar_res[EPI_WARPS];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *ring = smem + p.wbytes;
    float *tb = reinterpret_cast<float *>(ring + (size_t)p.nst * STAGE);
    float *rowscale = tb + p.tb_bytes / 4;                        // [SCALE_SLOTS][2][ROWS]
    float *bias_s = rowscale + SCALE_SLOTS * 2 * ROWS;            // [256] bias, then [256] inverse weight scale per column
    float *winv_s = bias_s + 256;
    uint8_t *raw = reinterpret_cast<uint8_t *>(winv_s + 256);     // STAGED: [NRAW][RAW_ROWS][K] fp32

    if (tid == 0) {
        for (int s = 0; s < p.nst; ++s) { mbar_init(&bar_full[s], LOAD_THREADS); mbar_init(&bar_empty[s], 1); }
        if (STAGED)
            for (int s = 0; s < NRAW; ++s) { mbar_init(&bar_raw_full[s], 1); mbar_init(&bar_raw_empty[s], LOAD_WARPS); }
        if (TMA_OUT)
            for (int w = 0; w < EPI_WARPS; ++w) mbar_init(&bar_res[w], 1);
        for (int b = 0; b < 2; ++b) { mbar_init(&bar_acc_full[b], 1); mbar_init(&bar_acc_empty[b], EPI_WARPS * 32); }
        mbar_init(&bar_w, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc<512>(&tmem_base_s);
    for (int i = tid; i < 256; i += (int)blockDim.x) {
        bias_s[i] = (p.bias != nullptr && i < p.N) ? p.bias[i] : 0.0f;
        winv_s[i] = i < p.Np ? reinterpret_cast<const float *>(p.packed + p.wbytes)[i] : 0.0f;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int spt = p.Kp / KC;                                    // stages per tile
    const int spg = p.gw / KC;                                    // stages per group
    const int acc_cols = p.groups * p.Np;                         // tensor-memory columns of one accumulator buffer

    if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        if (lane == 0) {
            mbar_expect_tx(&bar_w, (uint32_t)p.wbytes);
            for (int off = 0; off < p.wbytes; off += 16384) {
                const int n = p.wbytes - off < 16384 ? p.wbytes - off : 16384;
                bulk_g2s(smem + off, p.packed + off, (uint32_t)n, &bar_w);
            }
        }
        __syncwarp();
        const uint32_t idesc = idesc_f16(128, p.Np);
        const uint32_t sbase = smem_u32(smem), rbase = smem_u32(ring);
        const uint32_t w_sbo = (uint32_t)p.Kp * 16u, w_half = (uint32_t)p.Np * (uint32_t)p.Kp * 2u;
        mbar_wait(&bar_w, 0, p.status, 1);
        uint32_t sc = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            mbar_wait(&bar_acc_empty[b], (uint32_t)(((it >> 1) & 1) ^ 1), p.status, 2);
            tc_fence_after();
            for (int j = 0; j < spt; ++j, ++sc) {
                const uint32_t s = sc % (uint32_t)p.nst, ph = (sc / (uint32_t)p.nst) & 1u;
                mbar_wait(&bar_full[s], ph, p.status, 3);
                tc_fence_after();
                const int g = j / spg;
                const uint32_t dcol = tmem + (uint32_t)(b * acc_cols + g * p.Np);
                const uint32_t a_hi = rbase + s * STAGE, a_lo = a_hi + HALF;
                const uint32_t w_hi = sbase + (uint32_t)j * (KC / 8) * 128u, w_lo = w_hi + w_half;
                if (elect_one()) {
#pragma unroll
                    for (uint32_t kk = 0; kk < KC / 16; ++kk) {
                        const uint64_t ah = smem_desc(a_hi + kk * 2u * A_LBO, A_LBO, A_SBO);
                        const uint64_t al = smem_desc(a_lo + kk * 2u * A_LBO, A_LBO, A_SBO);
                        const uint64_t bh = smem_desc(w_hi + kk * 256u, 128u, w_sbo);
                        const uint64_t bl = smem_desc(w_lo + kk * 256u, 128u, w_sbo);
                        umma_f16(dcol, ah, bh, idesc, (j % spg) != 0 || kk != 0);
                        umma_f16(dcol, ah, bl, idesc, true);
                        umma_f16(dcol, al, bh, idesc, true);
                    }
                    umma_commit(&bar_empty[s]);
                    if (j == spt - 1) umma_commit(&bar_acc_full[b]);
                }
                __syncwarp();
            }
        }
    } else if (STAGED && warp == PROD_WARP) {
        // ================= producer (staged form): raw fp32 rows of x -> ring of RAW_ROWS-row chunks ==================
        // one tensor-map load per chunk (box = RAW_ROWS rows x K columns, dense rows in shared memory): any row stride costs the
        // same single instruction, rows past the matrix arrive as zeros.  (One bulk copy per ROW measured 61 us against 35 us
        // for the strided operands of the update: the copies of a chunk were issued one after the other.)
        const uint32_t chunk_bytes = RAW_ROWS * (uint32_t)p.K * 4u;
        uint32_t rs = 0, rph = 0;                                 // ring slot and its phase (the ring has 3 or 4 chunks)
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x) {
            for (int c = 0; c < ROWS / RAW_ROWS; ++c, rph ^= (rs + 1 == (uint32_t)p.nraw), rs = (rs + 1 == (uint32_t)p.nraw) ? 0u : rs + 1) {
                mbar_wait(&bar_raw_empty[rs], rph ^ 1u, p.status, 8);
                const long long row0 = (long long)tile * ROWS + c * RAW_ROWS;
                if (lane == 0) {
                    if (row0 < p.rows) {
                        mbar_expect_tx(&bar_raw_full[rs], chunk_bytes);
                        tma_load_2d(&xmap, smem_u32(raw + rs * chunk_bytes), 0, (int)row0, &bar_raw_full[rs]);
                    } else {
                        mbar_arrive(&bar_raw_full[rs]);            // a chunk entirely past the matrix: nothing to fetch
                    }
                }
                __syncwarp();
            }
        }
    } else if (STAGED && warp >= LOAD_WARP0) {
        // ================= converters (staged form) =================
        // lane = (row of the pass sub, 8-column chunk kc) as in the register loaders below; a chunk of the raw ring is one
        // pass of the eight warps (warp lw converts rows 4 lw .. 4 lw + 3 of the chunk).  One group (K = gw <= 128).
        const int lw = warp - LOAD_WARP0, sub = lane >> 3, kc = lane & 7;
        const int halves = p.gw / KC;
        const uint32_t row_bytes = (uint32_t)p.K * 4u, chunk_bytes = RAW_ROWS * row_bytes;
        uint32_t sc = 0, rs = 0, rph = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const uint32_t s0 = sc % (uint32_t)p.nst, ph0 = (sc / (uint32_t)p.nst) & 1u;
            const uint32_t s1 = (sc + 1) % (uint32_t)p.nst, ph1 = ((sc + 1) / (uint32_t)p.nst) & 1u;
            for (int c = 0; c < ROWS / RAW_ROWS; ++c, rph ^= (rs + 1 == (uint32_t)p.nraw), rs = (rs + 1 == (uint32_t)p.nraw) ? 0u : rs + 1) {
                const int r = c * RAW_ROWS + lw * 4 + sub;
                const bool live = (long long)tile * ROWS + r < p.rows;
                mbar_wait(&bar_raw_full[rs], rph, p.status, 7);
                const uint8_t *src = raw + rs * chunk_bytes + (uint32_t)(lw * 4 + sub) * row_bytes + (uint32_t)kc * 32u;
                float4 v[2][2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    v[h][0] = v[h][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (h < halves && live) {
                        v[h][0] = *reinterpret_cast<const float4 *>(src + h * 256);
                        v[h][1] = *reinterpret_cast<const float4 *>(src + h * 256 + 16);
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_raw_empty[rs]);       // the chunk's values are in registers: hand it back
                float m = fmaxf(fmaxf(amax4(v[0][0]), amax4(v[0][1])), fmaxf(amax4(v[1][0]), amax4(v[1][1])));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                float inv;
                const float s = pow2_scale(m, &inv);
                uint4 hi[2], lo[2];
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    scale4(v[h][0], s); scale4(v[h][1], s);
                    split8_f16(v[h][0], v[h][1], hi[h], lo[h]);
                }
                if (c == 0) {
                    // row-scale slot it & 1 may still be read by the epilogue of tile it - 2; the operand stages by the MMAs
                    mbar_wait(&bar_acc_empty[it & 1], (uint32_t)(((it >> 1) & 1) ^ 1), p.status, 6);
                    mbar_wait(&bar_empty[s0], ph0 ^ 1u, p.status, 4);
                    if (halves > 1) mbar_wait(&bar_empty[s1], ph1 ^ 1u, p.status, 4);
                }
                if (kc == 0) rowscale[((it & (SCALE_SLOTS - 1)) * 2) * ROWS + r] = inv;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    if (h < halves) {
                        uint8_t *st = ring + (size_t)(h == 0 ? s0 : s1) * STAGE + (uint32_t)kc * A_LBO + (uint32_t)r * 16u;
                        *reinterpret_cast<uint4 *>(st) = hi[h];
                        *reinterpret_cast<uint4 *>(st + HALF) = lo[h];
                    }
                }
            }
            fence_async_smem();
            mbar_arrive(&bar_full[s0]);
            if (halves > 1) mbar_arrive(&bar_full[s1]);
            sc += (uint32_t)halves;
        }
    } else if (warp >= LOAD_WARP0) {
        // ================= loaders =================
        // lane = (row-in-pass sub, 8-column chunk kc): a warp reads 4 rows x 256 contiguous bytes per pass.  A group (<= 128
        // columns of a 128-row tile) is processed in two batches of two passes (32 data registers in flight per thread);
        // both ring stages of the group are acquired first and released together.
        const int lw = warp - LOAD_WARP0, sub = lane >> 3, kc = lane & 7;
        const int halves = p.gw / KC;                             // 64-column halves of a group: 1 or 2
        uint32_t sc = 0;
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            for (int g = 0; g < p.groups; ++g) {
                const uint32_t s0 = sc % (uint32_t)p.nst, ph0 = (sc / (uint32_t)p.nst) & 1u;
                const uint32_t s1 = (sc + 1) % (uint32_t)p.nst, ph1 = ((sc + 1) / (uint32_t)p.nst) & 1u;
#pragma unroll
                for (int batch = 0; batch < NBATCH; ++batch) {
                    float4 v[2][2][2];
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const long long row = (long long)tile * ROWS + lw * RPW + (batch * 2 + q) * 4 + sub;
                        const float *src = p.x + row * p.ldx;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            const int col = g * p.gw + h * KC + kc * 8;
                            float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
                            if (h < halves && row < p.rows) {
                                if (p.vec_ok && col + 8 <= p.K) {
                                    a = __ldg(reinterpret_cast<const float4 *>(src + col));
                                    c = __ldg(reinterpret_cast<const float4 *>(src + col + 4));
                                } else if (col < p.K) {
                                    float t[8];
#pragma unroll
                                    for (int e = 0; e < 8; ++e) t[e] = col + e < p.K ? __ldg(src + col + e) : 0.0f;
                                    a = make_float4(t[0], t[1], t[2], t[3]);
                                    c = make_float4(t[4], t[5], t[6], t[7]);
                                }
                            }
                            v[q][h][0] = a; v[q][h][1] = c;
                        }
                    }
                    if (batch == 0) {
                        // the row scales of tile `it` live in slot it & 1, which the epilogue of tile it - 2 may still be reading:
                        // wait for it exactly as the MMA issuer does (the loads above are already in flight)
                        if (g == 0) mbar_wait(&bar_acc_empty[it & 1], (uint32_t)(((it >> 1) & 1) ^ 1), p.status, 6);
                        mbar_wait(&bar_empty[s0], ph0 ^ 1u, p.status, 4);
                        if (halves > 1) mbar_wait(&bar_empty[s1], ph1 ^ 1u, p.status, 4);
                    }
#pragma unroll
                    for (int q = 0; q < 2; ++q) {
                        const int r = lw * RPW + (batch * 2 + q) * 4 + sub;
                        float m = fmaxf(fmaxf(amax4(v[q][0][0]), amax4(v[q][0][1])), fmaxf(amax4(v[q][1][0]), amax4(v[q][1][1])));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 1));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 2));
                        m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, 4));
                        float inv;
                        const float s = pow2_scale(m, &inv);
                        if (kc == 0) rowscale[((it & (SCALE_SLOTS - 1)) * 2 + g) * ROWS + r] = inv;
#pragma unroll
                        for (int h = 0; h < 2; ++h) {
                            if (h < halves) {
                                scale4(v[q][h][0], s); scale4(v[q][h][1], s);
                                uint4 hi, lo;
                                split8_f16(v[q][h][0], v[q][h][1], hi, lo);
                                uint8_t *st = ring + (size_t)(h == 0 ? s0 : s1) * STAGE + (uint32_t)kc * A_LBO + (uint32_t)r * 16u;
                                *reinterpret_cast<uint4 *>(st) = hi;
                                *reinterpret_cast<uint4 *>(st + HALF) = lo;
                            }
                        }
                    }
                }
                fence_async_smem();
                mbar_arrive(&bar_full[s0]);
                if (halves > 1) mbar_arrive(&bar_full[s1]);
                sc += (uint32_t)halves;
            }
        }
    } else {
        // ================= epilogue: thread = accumulator row (phase A), 4 columns x 4 row octets (phase B) ==========
        // warps w and w + 4 share tensor-memory lanes 32 (w % 4) .. + 31 and split the 16-column blocks between them.
        // Phase A: tcgen05.ld of 16 columns of the thread's row, times the row's inverse scale, into the warp's 32 x 16
        // transposition block.  Phase B: lane = (row octet member l & 7, column quad l >> 3): 16-byte loads from the block,
        // times the column's inverse weight scale, + bias, ReLU, (+ previous output), 16-byte global stores -- one warp
        // store covers 8 rows x 64 contiguous bytes.
        const int lw = warp & 3, half = warp >> 2;
        const int r = lw * 32 + lane;
        const uint32_t trow = tmem + ((uint32_t)(lw * 32) << 16);
        float *tbw = tb + warp * 32 * TB_STRIDE;
        const int nblk = (p.N + 15) / 16;                              // 16-column blocks of the output
        const int rq = lane & 7, cq = lane >> 3;                       // phase-B row (within an octet) and column quad
        const bool vec_out = (p.ldo % 4 == 0) && (((uintptr_t)p.out & 15) == 0);
        const bool vec_res = (p.ldr % 4 == 0) && (((uintptr_t)p.res & 15) == 0);
        const bool two_groups = !STAGED && p.groups > 1;               // (the staged form has one scale group: K <= 128)
        uint32_t res_ph = 0;
        if (TMA_OUT && p.acc && lane == 0 && (int)blockIdx.x < p.n_tiles && half < nblk) {      // first block of the added matrix
            const uint32_t tb_a = smem_u32(tb);
            mbar_expect_tx(&bar_res[warp], 2048u);
            tma_load_2d(&rmap, ((tb_a + 511u) & ~511u) + (EPI_WARPS + warp) * 2048u, half * 16, (int)blockIdx.x * ROWS + lw * 32, &bar_res[warp]);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < p.n_tiles; tile += gridDim.x, ++it) {
            const int b = it & 1;
            const long long row0 = (long long)tile * ROWS + lw * 32;
            // accumulate mode: the previous output does not depend on the MMAs -- fetch the first block's values early
            mbar_wait(&bar_acc_full[b], (uint32_t)((it >> 1) & 1), p.status, 5);
            tc_fence_after();
            const float *rs = rowscale + (it & (SCALE_SLOTS - 1)) * 2 * ROWS;
            const float inv0 = rs[r], inv1 = p.groups > 1 ? rs[ROWS + r] : 0.0f;
            if constexpr (TMA_OUT) {
                // thread = row: scale, bias, ReLU in registers, 4 x STS.128 into the warp's swizzled 32 x 16 tile, one tensor store
                // (offsets from the shared array, not a rounded pointer: the compiler must still see shared-memory accesses)
                const uint32_t tb_a = smem_u32(tb), tile_off = (uint32_t)(reinterpret_cast<uint8_t *>(tb) - smem) + (((tb_a + 511u) & ~511u) - tb_a);
                uint8_t *tile_s = smem + tile_off + warp * 2048;
                const uint32_t tile_a = smem_u32(tile_s), sw = (uint32_t)((lane >> 1) & 3);
                // accumulate mode: the 32 x 16 block of the added matrix arrives by a tensor-map LOAD into a second swizzled tile
                // (one block ahead, across tile boundaries), and the row thread reads its 64 bytes from there
                const uint8_t *res_s = tile_s + EPI_WARPS * 2048;
                for (int c = half; c < nblk; c += 2) {
                    uint32_t v0[16], v1[16];
                    tmem_ld16(trow + (uint32_t)(b * acc_cols + c * 16), v0);
                    if (two_groups) tmem_ld16(trow + (uint32_t)(b * acc_cols + p.Np + c * 16), v1);
                    float4 rs[4];
                    if (p.acc) {
                        mbar_wait(&bar_res[warp], res_ph, p.status, 9);
                        res_ph ^= 1u;
#pragma unroll
                        for (int j = 0; j < 4; ++j) rs[j] = *reinterpret_cast<const float4 *>(res_s + lane * 64 + (((uint32_t)j ^ sw) << 4));
                        __syncwarp();                                  // every lane has its values: the tile may be refilled
                        int tn = tile, cn = c + 2;
                        if (cn >= nblk) { tn += (int)gridDim.x; cn = half; }
                        if (lane == 0 && tn < p.n_tiles && cn < nblk) {
                            mbar_expect_tx(&bar_res[warp], 2048u);
                            tma_load_2d(&rmap, smem_u32(res_s), cn * 16, tn * ROWS + lw * 32, &bar_res[warp]);
                        }
                    }
                    tmem_ld_wait();
                    if (lane == 0) tma_store_wait_read();             // the previous block's store has read the tile
                    __syncwarp();
#pragma unroll
                    for (int j = 0; j < 16; j += 4) {
                        const float4 wi = *reinterpret_cast<const float4 *>(winv_s + c * 16 + j);
                        const float4 bi = *reinterpret_cast<const float4 *>(bias_s + c * 16 + j);
                        float4 f;
                        f.x = __uint_as_float(v0[j]) * inv0; f.y = __uint_as_float(v0[j + 1]) * inv0;
                        f.z = __uint_as_float(v0[j + 2]) * inv0; f.w = __uint_as_float(v0[j + 3]) * inv0;
                        if (two_groups) {
                            f.x = fmaf(__uint_as_float(v1[j]), inv1, f.x); f.y = fmaf(__uint_as_float(v1[j + 1]), inv1, f.y);
                            f.z = fmaf(__uint_as_float(v1[j + 2]), inv1, f.z); f.w = fmaf(__uint_as_float(v1[j + 3]), inv1, f.w);
                        }
                        f.x = fmaf(f.x, wi.x, bi.x); f.y = fmaf(f.y, wi.y, bi.y); f.z = fmaf(f.z, wi.z, bi.z); f.w = fmaf(f.w, wi.w, bi.w);
                        if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
                        if (p.acc) { f.x += rs[j >> 2].x; f.y += rs[j >> 2].y; f.z += rs[j >> 2].z; f.w += rs[j >> 2].w; }
                        *reinterpret_cast<float4 *>(tile_s + lane * 64 + ((((uint32_t)j >> 2) ^ sw) << 4)) = f;
                    }
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) tma_store_2d(&omap, tile_a, c * 16, (int)(row0));
                }
            } else
            for (int c = half; c < nblk; c += 2) {
                const int col = c * 16 + cq * 4;
                float4 old[4];
                if (p.acc) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const long long row = row0 + i * 8 + rq;
                        old[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (row < p.rows && col < p.N) {
                            const float *src = p.res + row * p.ldr + col;
                            if (vec_res && col + 4 <= p.N) old[i] = *reinterpret_cast<const float4 *>(src);
                            else {
                                old[i].x = src[0];
                                if (col + 1 < p.N) old[i].y = src[1];
                                if (col + 2 < p.N) old[i].z = src[2];
                                if (col + 3 < p.N) old[i].w = src[3];
                            }
                        }
                    }
                }
                uint32_t v0[16], v1[16];
                tmem_ld16(trow + (uint32_t)(b * acc_cols + c * 16), v0);
                if (p.groups > 1) tmem_ld16(trow + (uint32_t)(b * acc_cols + p.Np + c * 16), v1);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 16; j += 4) {
                    float4 f;
                    f.x = __uint_as_float(v0[j]) * inv0; f.y = __uint_as_float(v0[j + 1]) * inv0;
                    f.z = __uint_as_float(v0[j + 2]) * inv0; f.w = __uint_as_float(v0[j + 3]) * inv0;
                    if (p.groups > 1) {
                        f.x = fmaf(__uint_as_float(v1[j]), inv1, f.x); f.y = fmaf(__uint_as_float(v1[j + 1]), inv1, f.y);
                        f.z = fmaf(__uint_as_float(v1[j + 2]), inv1, f.z); f.w = fmaf(__uint_as_float(v1[j + 3]), inv1, f.w);
                    }
                    *reinterpret_cast<float4 *>(tbw + lane * TB_STRIDE + j) = f;
                }
                __syncwarp();
                const float4 wi = *reinterpret_cast<const float4 *>(winv_s + col);
                const float4 bi = *reinterpret_cast<const float4 *>(bias_s + col);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rr = i * 8 + rq;
                    const long long row = row0 + rr;
                    float4 f = *reinterpret_cast<const float4 *>(tbw + rr * TB_STRIDE + cq * 4);
                    f.x = fmaf(f.x, wi.x, bi.x); f.y = fmaf(f.y, wi.y, bi.y); f.z = fmaf(f.z, wi.z, bi.z); f.w = fmaf(f.w, wi.w, bi.w);
                    if (p.relu) { f.x = fmaxf(f.x, 0.f); f.y = fmaxf(f.y, 0.f); f.z = fmaxf(f.z, 0.f); f.w = fmaxf(f.w, 0.f); }
                    if (p.acc) { f.x += old[i].x; f.y += old[i].y; f.z += old[i].z; f.w += old[i].w; }
                    if (row < p.rows && col < p.N) {
                        float *dst = p.out + row * p.ldo + col;
                        if (vec_out && col + 4 <= p.N) *reinterpret_cast<float4 *>(dst) = f;
                        else {
                            dst[0] = f.x;
                            if (col + 1 < p.N) dst[1] = f.y;
                            if (col + 2 < p.N) dst[2] = f.z;
                            if (col + 3 < p.N) dst[3] = f.w;
                        }
                    }
                }
                __syncwarp();
            }
            tc_fence_before();
            mbar_arrive(&bar_acc_empty[b]);
        }
    }
    if (TMA_OUT && warp < EPI_WARPS && lane == 0) tma_store_wait_all();
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<512>(tmem);
}

// ---- weight packing ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) tg_pack_kernel(const float *w, int N, int K, int Np, int Kp, int ld, int transposed,
                                                      uint8_t *out) {
    // one CTA = 8 consecutive rows n of B (one row group of core matrices); one warp = one row: its own power-of-two
    // scale (the epilogue multiplies column n of the product by the inverse), then the hi / lo planes of that row
    const int n = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    float m = 0.0f;
    if (n < N)
        for (int k = lane; k < K; k += 32) m = fmaxf(m, fabsf(w[transposed ? (size_t)k * ld + n : (size_t)n * ld + k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    float inv;
    const float s = pow2_scale(m, &inv);
    if (lane == 0) reinterpret_cast<float *>(out + (size_t)4 * Np * Kp)[n] = n < N ? inv : 0.0f;
    __half *hi = reinterpret_cast<__half *>(out), *lo = hi + (size_t)Np * Kp;
    // element (n, k) of the canonical K-major order: ((n/8) * (Kp/8) + k/8) * 64 + (n%8) * 8 + k%8
    for (int k = lane; k < Kp; k += 32) {
        float x = 0.0f;
        if (n < N && k < K) x = w[transposed ? (size_t)k * ld + n : (size_t)n * ld + k] * s;
        const __half h = __float2half_rn(x);
        const size_t i = ((size_t)(n >> 3) * (Kp >> 3) + (k >> 3)) * 64 + (size_t)(n & 7) * 8 + (k & 7);
        hi[i] = h;
        lo[i] = __float2half_rn(x - __half2float(h));
    }
}

// ---- weight gradient: dW = x^T y ----------------------------------------------------------------------------------
// Row step = 64 rows.  An operand block = 64 rows x 128 features of x or y as three bf16 planes (h, m, l), laid out like
// the tiles above (feature chunk kc at kc * FB_LBO, row r at + r * 16).  Read as an MN-major operand: mn = feature
// (M = N = 128), k = row: 8 features of one row are the 16 contiguous bytes, 8 rows the 128-byte core matrix,
// SBO = FB_LBO (next 8 features), LBO = 128 (next 8 rows); one MMA consumes 16 rows (+256 bytes).
// Two block heights: 64 rows (ring of 4 blocks: two row steps of a one-block-per-operand product in flight) and 32 rows
// (ring of 8) for products with a 256-wide operand, whose row step needs THREE blocks -- with 64-row blocks only one more
// block fits beside a step, and the loaders idle while its twelve-MMA groups run.
template <int WR, bool STAGED = false> struct WG {
    static constexpr uint32_t LBO = WR * 16 + 16, PLANE = 16 * LBO, BLOCK = 3 * PLANE;      // 64: 1040, 16640, 49920; 32: 528, 8448, 25344
    // staged form: the raw fp32 rows of an operand block (WR rows x <= 128 columns) arrive in a ring of 64 KB by tensor-map loads
    // and the loader warps convert from there; the operand ring gives up a quarter of its blocks for it (conversion, not the
    // MMAs, is what a row step waits for, so three / six blocks keep the tensor pipe fed)
    static constexpr int NB = STAGED ? (WR == 64 ? 3 : 6) : (WR == 64 ? 4 : 8);
    static constexpr uint32_t RAW_CHUNK = WR * 512u;
    static constexpr int NRAW = STAGED ? 65536 / (int)RAW_CHUNK : 0;
    static constexpr int SMEM = NB * (int)BLOCK + NRAW * (int)RAW_CHUNK;
};
constexpr int MAX_NB = 8, MAX_WRAW = 4;

struct WgParams {
    const float *x, *y;
    float *partial;
    uint32_t *status;
    long long rows;
    int ldx, ldy, a, b, ab, bb, n_steps, vx, vy;
    uint32_t lbo, sbo;
};

__host__ __device__ constexpr uint32_t idesc_bf16_mn(int M, int N) {
    return (1u << 4) /* D = f32 */ | (1u << 7) /* A = bf16 */ | (1u << 10) /* B = bf16 */ | (1u << 15) /* A MN-major */ |
           (1u << 16) /* B MN-major */ | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

__device__ __forceinline__ uint32_t pack_bf2(float a, float b) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_bf2(uint32_t u) { return make_float2(__uint_as_float(u << 16), __uint_as_float(u & 0xffff0000u)); }

__device__ __forceinline__ void split8_bf16(const float4 &a, const float4 &b, uint4 &h, uint4 &m, uint4 &l) {
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    uint32_t hh[4], mm[4], ll[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        hh[j] = pack_bf2(v[2 * j], v[2 * j + 1]);
        const float2 f = unpack_bf2(hh[j]);
        const float r0 = v[2 * j] - f.x, r1 = v[2 * j + 1] - f.y;
        mm[j] = pack_bf2(r0, r1);
        const float2 g = unpack_bf2(mm[j]);
        ll[j] = pack_bf2(r0 - g.x, r1 - g.y);
    }
    h = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    m = make_uint4(mm[0], mm[1], mm[2], mm[3]);
    l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
}

template <int WROWS, bool STAGED>
__global__ void __launch_bounds__(STAGED ? THREADS_STAGED : THREADS, 1)
tg_wgrad_kernel(const WgParams p, const __grid_constant__ CUtensorMap xmap, const __grid_constant__ CUtensorMap ymap) {
    typedef WG<WROWS, STAGED> L;
    constexpr uint32_t FB_LBO = L::LBO, PLANE = L::PLANE, BLOCK = L::BLOCK;
    constexpr int NB = L::NB;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_full[MAX_NB], bar_empty[MAX_NB], bar_done, bar_raw_full[MAX_WRAW], bar_raw_empty[MAX_WRAW];
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    uint8_t *raw = smem + (size_t)NB * BLOCK;                      // STAGED: [NRAW][WROWS][<= 128] fp32, dense rows
    if (tid == 0) {
        for (int s = 0; s < NB; ++s) { mbar_init(&bar_full[s], LOAD_THREADS); mbar_init(&bar_empty[s], 1); }
        for (int s = 0; s < L::NRAW; ++s) { mbar_init(&bar_raw_full[s], 1); mbar_init(&bar_raw_empty[s], LOAD_WARPS); }
        mbar_init(&bar_done, 1);
        mbar_fence_init();
    }
    if (warp == MMA_WARP) tmem_alloc<512>(&tmem_base_s);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;
    const int nblk = p.ab + p.bb;                                 // operand blocks per row step

    if (warp == MMA_WARP) {
        const uint32_t idesc = idesc_bf16_mn(128, 128);
        const uint32_t sbase = smem_u32(smem);
        uint32_t bc = 0;
        bool first = true;
        for (int step = blockIdx.x; step < p.n_steps; step += gridDim.x) {
            uint32_t slot[4];
            for (int k = 0; k < nblk; ++k, ++bc) {
                slot[k] = bc % NB;
                mbar_wait(&bar_full[slot[k]], (bc / NB) & 1u, p.status, 11);
            }
            tc_fence_after();
            if (elect_one()) {
                for (int i = 0; i < p.ab; ++i)
                    for (int j = 0; j < p.bb; ++j) {
                        const uint32_t dcol = tmem + (uint32_t)(i * p.bb + j) * 128u;
                        const uint32_t xa = sbase + slot[i] * BLOCK, ya = sbase + slot[p.ab + j] * BLOCK;
#pragma unroll
                        for (uint32_t kk = 0; kk < WROWS / 16; ++kk) {
                            uint64_t xd[3], yd[3];
#pragma unroll
                            for (uint32_t c = 0; c < 3; ++c) {
                                xd[c] = smem_desc(xa + c * PLANE + kk * 256u, p.lbo, p.sbo);
                                yd[c] = smem_desc(ya + c * PLANE + kk * 256u, p.lbo, p.sbo);
                            }
                            umma_f16(dcol, xd[0], yd[0], idesc, !first || kk != 0);
                            umma_f16(dcol, xd[0], yd[1], idesc, true);
                            umma_f16(dcol, xd[1], yd[0], idesc, true);
                            umma_f16(dcol, xd[0], yd[2], idesc, true);
                            umma_f16(dcol, xd[2], yd[0], idesc, true);
                            umma_f16(dcol, xd[1], yd[1], idesc, true);
                        }
                    }
                for (int k = 0; k < nblk; ++k) umma_commit(&bar_empty[slot[k]]);
            }
            __syncwarp();
            first = false;
        }
        if (elect_one()) umma_commit(&bar_done);
        __syncwarp();
    } else if (STAGED && warp == PROD_WARP) {
        // producer (staged form): one tensor-map load per operand block (WROWS rows x the block's columns, dense rows; rows past
        // the matrix arrive as zeros) into the raw ring
        if (lane == 0) {
            uint32_t rc = 0;
            for (int step = blockIdx.x; step < p.n_steps; step += gridDim.x) {
                for (int k = 0; k < nblk; ++k, ++rc) {
                    const uint32_t rs = rc % (uint32_t)L::NRAW, rph = (rc / (uint32_t)L::NRAW) & 1u;
                    const bool isx = k < p.ab;
                    const int width = isx ? p.a : p.b, f0 = (isx ? k : k - p.ab) * 128, bw = width < 128 ? width : 128;
                    mbar_wait(&bar_raw_empty[rs], rph ^ 1u, p.status, 14);
                    mbar_expect_tx(&bar_raw_full[rs], (uint32_t)(WROWS * bw * 4));
                    tma_load_2d(isx ? &xmap : &ymap, smem_u32(raw + rs * L::RAW_CHUNK), f0, step * WROWS, &bar_raw_full[rs]);
                }
            }
        }
        __syncwarp();
    } else if (STAGED && warp >= LOAD_WARP0) {
        // converters (staged form): the same lane mapping and split as the register loaders below, rows read from the raw ring
        const int lw = warp - LOAD_WARP0, sub = lane >> 3, kc = lane & 7;
        constexpr int WQ = WROWS / LOAD_WARPS / 4;
        uint32_t bc = 0;
        for (int step = blockIdx.x; step < p.n_steps; step += gridDim.x) {
            for (int k = 0; k < nblk; ++k, ++bc) {
                const uint32_t rs = bc % (uint32_t)L::NRAW, rph = (bc / (uint32_t)L::NRAW) & 1u, s = bc % NB;
                const bool isx = k < p.ab;
                const int width = isx ? p.a : p.b, bw = width < 128 ? width : 128;
                mbar_wait(&bar_raw_full[rs], rph, p.status, 15);
                float4 v[WQ][2][2];
#pragma unroll
                for (int q = 0; q < WQ; ++q) {
                    const int r = lw * (WQ * 4) + q * 4 + sub;
                    const uint8_t *src = raw + rs * L::RAW_CHUNK + (uint32_t)r * (uint32_t)(bw * 4) + (uint32_t)kc * 32u;
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        v[q][h][0] = v[q][h][1] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (h * 64 + kc * 8 < bw) {
                            v[q][h][0] = *reinterpret_cast<const float4 *>(src + h * 256);
                            v[q][h][1] = *reinterpret_cast<const float4 *>(src + h * 256 + 16);
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_raw_empty[rs]);
                mbar_wait(&bar_empty[s], ((bc / NB) & 1u) ^ 1u, p.status, 12);
                uint8_t *blk = smem + (size_t)s * BLOCK;
#pragma unroll
                for (int q = 0; q < WQ; ++q)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint4 hh, mm, ll;
                        split8_bf16(v[q][h][0], v[q][h][1], hh, mm, ll);
                        uint8_t *d = blk + (uint32_t)(h * 8 + kc) * FB_LBO + (uint32_t)(lw * (WQ * 4) + q * 4 + sub) * 16u;
                        *reinterpret_cast<uint4 *>(d) = hh;
                        *reinterpret_cast<uint4 *>(d + PLANE) = mm;
                        *reinterpret_cast<uint4 *>(d + 2 * PLANE) = ll;
                    }
                fence_async_smem();
                mbar_arrive(&bar_full[s]);
            }
        }
    } else if (warp >= LOAD_WARP0) {
        // loaders: the global loads of operand block idx + 1 are in flight while block idx is split and stored (two register
        // buffers), so the load latency is paid once per CTA, not once per block
        const int lw = warp - LOAD_WARP0, sub = lane >> 3, kc = lane & 7;
        const int my_steps = (p.n_steps - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
        const int total = my_steps * nblk;
        constexpr int WQ = WROWS / LOAD_WARPS / 4;                 // passes of 4 rows per loader warp and operand block
        auto load_block = [&](int idx, float4 (&v)[WQ][2][2]) {
            const int step = (int)blockIdx.x + (idx / nblk) * (int)gridDim.x, k = idx % nblk;
            const bool isx = k < p.ab;
            const float *src = isx ? p.x : p.y;
            const int ld = isx ? p.ldx : p.ldy, width = isx ? p.a : p.b, f0 = (isx ? k : k - p.ab) * 128;
            const int vec = isx ? p.vx : p.vy;
#pragma unroll
            for (int q = 0; q < WQ; ++q) {
                const long long row = (long long)step * WROWS + lw * (WQ * 4) + q * 4 + sub;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int col = f0 + h * 64 + kc * 8;
                    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), c = a;
                    if (idx < total && row < p.rows && col < width) {
                        const float *s = src + row * ld + col;
                        if (vec && col + 8 <= width) {
                            a = __ldg(reinterpret_cast<const float4 *>(s));
                            c = __ldg(reinterpret_cast<const float4 *>(s + 4));
                        } else {
                            float t[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) t[e] = col + e < width ? __ldg(s + e) : 0.0f;
                            a = make_float4(t[0], t[1], t[2], t[3]);
                            c = make_float4(t[4], t[5], t[6], t[7]);
                        }
                    }
                    v[q][h][0] = a; v[q][h][1] = c;
                }
            }
        };
        auto emit_block = [&](int idx, const float4 (&v)[WQ][2][2]) {
            const uint32_t bc = (uint32_t)idx, s = bc % NB;
            mbar_wait(&bar_empty[s], ((bc / NB) & 1u) ^ 1u, p.status, 12);
            uint8_t *blk = smem + (size_t)s * BLOCK;
#pragma unroll
            for (int q = 0; q < WQ; ++q)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint4 hh, mm, ll;
                    split8_bf16(v[q][h][0], v[q][h][1], hh, mm, ll);
                    uint8_t *d = blk + (uint32_t)(h * 8 + kc) * FB_LBO + (uint32_t)(lw * (WQ * 4) + q * 4 + sub) * 16u;
                    *reinterpret_cast<uint4 *>(d) = hh;
                    *reinterpret_cast<uint4 *>(d + PLANE) = mm;
                    *reinterpret_cast<uint4 *>(d + 2 * PLANE) = ll;
                }
            fence_async_smem();
            mbar_arrive(&bar_full[s]);
        };
        float4 va[WQ][2][2], vb[WQ][2][2];
        load_block(0, va);
        for (int idx = 0; idx < total; idx += 2) {
            load_block(idx + 1, vb);
            emit_block(idx, va);
            if (idx + 1 < total) {
                load_block(idx + 2, va);
                emit_block(idx + 1, vb);
            }
        }
    } else {
        // epilogue: one partial [a][b] block per CTA; thread = x feature (accumulator row); warps w and w + 4 share a lane
        // window and split the 32-column chunks
        mbar_wait(&bar_done, 0, p.status, 13);
        tc_fence_after();
        const int lw = warp & 3, half = warp >> 2;
        const uint32_t trow = tmem + ((uint32_t)(lw * 32) << 16);
        float *dst = p.partial + (size_t)blockIdx.x * p.a * p.b;
        for (int i = 0; i < p.ab; ++i) {
            const int fa = i * 128 + lw * 32 + lane;
            for (int j = 0; j < p.bb; ++j)
                for (int c = half; c < 4; c += 2) {
                    uint32_t v[32];
                    tmem_ld32(trow + (uint32_t)((i * p.bb + j) * 128 + c * 32), v);
                    tmem_ld_wait();
                    if (fa < p.a) {
#pragma unroll
                        for (int t = 0; t < 32; ++t) {
                            const int fb = j * 128 + c * 32 + t;
                            if (fb < p.b) dst[(size_t)fa * p.b + fb] = __uint_as_float(v[t]);
                        }
                    }
                }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == MMA_WARP) tmem_dealloc<512>(tmem);
}

__global__ void __launch_bounds__(256) tg_reduce_kernel(const float *partial, int n_part, int a, int b, float *out, int ldo, int accumulate) {
    // eight lanes per output element: lane q adds partials q, q + 8, ... (consecutive lanes of an octet read the SAME element
    // of eight different partial blocks; the four octets of a warp read four consecutive elements), then a three-step
    // butterfly in a fixed order -- bit-reproducible, no atomics, and 8 x the loads in flight of one thread per element
    const int t = blockIdx.x * 256 + threadIdx.x, i = t >> 3, q = t & 7;
    const bool live = i < a * b;
    const size_t stride = (size_t)a * b;
    float s0 = 0.0f, s1 = 0.0f;
    if (live) {
        int c = q;
        for (; c + 8 < n_part; c += 16) { s0 += partial[(size_t)c * stride + i]; s1 += partial[(size_t)(c + 8) * stride + i]; }
        if (c < n_part) s0 += partial[(size_t)c * stride + i];
    }
    float s = s0 + s1;
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (live && q == 0) {
        float *d = out + (size_t)(i / b) * ldo + i % b;
        *d = accumulate ? *d + s : s;
    }
}

// ---- optimizer step ---------------------------------------------------------------------------------------------
struct AdamParams {
    TgTensor t[TG_MAX_TENSORS];
    int n;
    float lr, b1, b2, eps, max_norm;
    const float *grad_scale;
    long long *step;
    float *scratch, *total_out;
};
constexpr int ADAM_BLOCKS = 128;
static_assert(ADAM_BLOCKS <= TG_ADAM_SCRATCH_FLOATS, "one partial per block");

__global__ void __launch_bounds__(256) tg_adam_norm_kernel(const AdamParams p) {
    __shared__ float red[8];
    const float gs = p.grad_scale != nullptr ? *p.grad_scale : 1.0f;
    float acc = 0.0f;
    for (int k = 0; k < p.n; ++k) {
        float *g = p.t[k].g;
        if (g == nullptr) continue;
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < p.t[k].numel; i += (long long)ADAM_BLOCKS * 256) {
            const float x = g[i] * gs;
            if (p.grad_scale != nullptr) g[i] = x;
            acc = fmaf(x, x, acc);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int i = 0; i < 8; ++i) s += red[i];
        p.scratch[blockIdx.x] = s;
        if (blockIdx.x == 0) *p.step += 1;
    }
}

__global__ void __launch_bounds__(256) tg_adam_update_kernel(const AdamParams p) {
    __shared__ float s_coef, s_c1, s_c2;
    if (threadIdx.x == 0) {
        float s = 0.0f;
        for (int i = 0; i < ADAM_BLOCKS; ++i) s += p.scratch[i];
        const float total = sqrtf(s);
        if (blockIdx.x == 0 && p.total_out != nullptr) *p.total_out = total;
        s_coef = p.max_norm > 0.0f ? fminf(1.0f, p.max_norm / (total + 1e-6f)) : 1.0f;
        const double st = (double)*p.step;
        s_c1 = (float)(1.0 - pow((double)p.b1, st));
        s_c2 = (float)sqrt(1.0 - pow((double)p.b2, st));
    }
    __syncthreads();
    const float coef = s_coef, step_size = p.lr / s_c1, c2 = s_c2;
    for (int k = 0; k < p.n; ++k) {
        const TgTensor t = p.t[k];
        if (t.g == nullptr) continue;
        for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < t.numel; i += (long long)gridDim.x * 256) {
            const float g = t.g[i] * coef;
            const float m = p.b1 * t.m[i] + (1.0f - p.b1) * g;
            const float v = p.b2 * t.v[i] + (1.0f - p.b2) * g * g;
            t.m[i] = m; t.v[i] = v;
            t.p[i] -= step_size * (m / (sqrtf(v) / c2 + p.eps));
        }
    }
}

}  // namespace tg

// ---- host side ----------------------------------------------------------------------------------------------------
namespace {
constexpr int MAX_DEVICES = 64;
bool g_ready[MAX_DEVICES] = {};
int g_sms[MAX_DEVICES] = {};
uint32_t g_wg_lbo = 128, g_wg_sbo = 0;      // descriptor fields of tg_wgrad's MN-major operands (sbo 0 = the block's feature-chunk stride)
int g_wg_rows = 0;             // tg_debug_wgrad_rows: force the block height of tg_wgrad (32 / 64; 0 = by shape)
bool g_wg_staged = true;       // tg_debug_wgrad_staged(0): keep tg_wgrad on the register loaders
int g_tma_out = 1;             // tg_debug_tma_out(0): keep tg_linear's epilogue on STG stores
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {     // the driver entry point, looked up once (no link-time dependency on libcuda)
    static EncodeTiledFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *ptr = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)ptr;
    }
    return fn;
}
bool g_staged = true;          // tg_debug_staged(0): keep tg_linear on the register loaders (A/B measurements, tests of both forms)

int prepare(int *sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return fa_internal_fail(-3, "tg: no usable CUDA device");
    if (!g_ready[dev]) {
        e = cudaFuncSetAttribute(tg::tg_linear_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_linear_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_linear_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_linear_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::SMEM_LIMIT);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_wgrad_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::WG<64, false>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_wgrad_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::WG<32, false>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_wgrad_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::WG<64, true>::SMEM);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(tg::tg_wgrad_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, tg::WG<32, true>::SMEM);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return fa_internal_fail(-2, "tg: device setup: %s", cudaGetErrorString(e));
        g_ready[dev] = true;
    }
    if (sms) *sms = g_sms[dev];
    return 0;
}
int pad_to(int v, int m) { return (v + m - 1) / m * m; }
}  // namespace

extern "C" size_t tg_packed_bytes(int N, int K) { return (size_t)4 * pad_to(N, 16) * pad_to(K, 64) + (size_t)4 * pad_to(N, 16); }

extern "C" int tg_pack_weight(const float *d_w, int N, int K, int ld, int transposed, void *d_packed, void *stream) {
    if (!d_w || !d_packed) return fa_internal_fail(-1, "tg_pack_weight: NULL pointer");
    if (N < 1 || N > 256 || K < 1 || K > 256 || pad_to(N, 16) * pad_to(K, 64) > 32768 || ld < (transposed ? N : K))
        return fa_internal_fail(-1, "tg_pack_weight: need 1 <= N, K <= 256, padded N * K <= 32768, ld >= row length (N=%d K=%d ld=%d)", N, K, ld);
    if ((uintptr_t)d_packed & 15) return fa_internal_fail(-4, "tg_pack_weight: d_packed must be 16-byte aligned");
    tg::tg_pack_kernel<<<pad_to(N, 16) / 8, 256, 0, (cudaStream_t)stream>>>(d_w, N, K, pad_to(N, 16), pad_to(K, 64), ld, transposed, (uint8_t *)d_packed);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_pack_weight: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int tg_linear(const float *d_x, int ldx, long long rows, int K, const void *d_packed, int N, const float *d_bias,
                         int relu, int accumulate, float *d_out, int ldo, uint32_t *d_status, void *stream) {
    return tg_linear_res(d_x, ldx, rows, K, d_packed, N, d_bias, relu, accumulate ? d_out : nullptr, ldo, d_out, ldo, d_status, stream);
}

extern "C" int tg_linear_res(const float *d_x, int ldx, long long rows, int K, const void *d_packed, int N, const float *d_bias,
                             int relu, const float *d_res, int ldr, float *d_out, int ldo, uint32_t *d_status, void *stream) {
    const int accumulate = d_res != nullptr;
    if (!d_x || !d_packed || !d_out || !d_status) return fa_internal_fail(-1, "tg_linear: NULL pointer");
    if (rows < 1 || K < 1 || K > 256 || N < 1 || N > 256 || ldx < K || ldo < N || (accumulate && ldr < N))
        return fa_internal_fail(-1, "tg_linear: need rows >= 1, 1 <= K, N <= 256, ldx >= K, ldo >= N, ldr >= N");
    const int Kp = pad_to(K, 64), Np = pad_to(N, 16);
    if (Kp * Np > 32768) return fa_internal_fail(-1, "tg_linear: padded N * K must be <= 32768 (got %d x %d)", Np, Kp);
    if ((uintptr_t)d_packed & 15) return fa_internal_fail(-4, "tg_linear: d_packed must be 16-byte aligned");
    int sms = 0;
    if (int rc = prepare(&sms)) return rc;
    tg::LinParams p = {};
    p.x = d_x; p.packed = (const uint8_t *)d_packed; p.bias = d_bias; p.out = d_out; p.status = d_status; p.rows = rows;
    p.res = d_res; p.ldr = accumulate ? ldr : ldo;
    p.ldx = ldx; p.ldo = ldo; p.K = K; p.Kp = Kp; p.N = N; p.Np = Np; p.relu = relu; p.acc = accumulate;
    p.n_tiles = (int)((rows + tg::ROWS - 1) / tg::ROWS);
    p.gw = Kp >= 128 ? 128 : 64;
    p.groups = Kp / p.gw;
    p.wbytes = 4 * Np * Kp;
    p.vec_ok = (ldx % 4 == 0) && (((uintptr_t)d_x & 15) == 0);
    // output through the TMA engine: at least one whole 16-column block, rows (of the output and of an accumulated term, which
    // then arrives by tensor-map loads) that are 16-byte multiples apart
    CUtensorMap omap, rmap;
    memset(&omap, 0, sizeof(omap));
    memset(&rmap, 0, sizeof(rmap));
    if (g_tma_out && N >= 16 && N % 4 == 0 && ldo % 4 == 0 && (((uintptr_t)d_out) & 15) == 0 &&
        (!accumulate || (ldr % 4 == 0 && (((uintptr_t)d_res) & 15) == 0)) && encode_tiled()) {
        const cuuint64_t gdim[2] = {(cuuint64_t)N, (cuuint64_t)rows}, gstride[1] = {(cuuint64_t)ldo * 4}, rstride[1] = {(cuuint64_t)ldr * 4};
        const cuuint32_t box[2] = {16, 32}, estr[2] = {1, 1};
        CUresult r = encode_tiled()(&omap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_out, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                   CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r == CUDA_SUCCESS && accumulate)
            r = encode_tiled()(&rmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_res), gdim, rstride, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        p.tma_out = r == CUDA_SUCCESS;
    }
    // epilogue tile area: the transposition blocks of the STG form, or 512-byte aligned 2 KB tiles per warp (output; + added term)
    p.tb_bytes = !p.tma_out ? tg::TB_BYTES : (512 + tg::EPI_WARPS * 2048 * (accumulate ? 2 : 1));
    if (p.tb_bytes < tg::TB_BYTES) p.tb_bytes = tg::TB_BYTES;
    if (tg::SMEM_LIMIT - (p.tb_bytes + tg::RS_BYTES + tg::BIAS_BYTES) - p.wbytes < 2 * (int)tg::STAGE) {
        p.tma_out = 0;                          // (256-wide weights + an added term: no room for the second tile set)
        p.tb_bytes = tg::TB_BYTES;
    }
    const int fixed = p.tb_bytes + tg::RS_BYTES + tg::BIAS_BYTES;
    // staged form: K = 64 or 128 exactly (one scale group, rows are whole 16-byte multiples), 16-byte aligned rows, and room
    // for a raw ring of 4 (or 3) chunks beside the weights and at least one tile of operand stages
    const int chunk = tg::RAW_ROWS * K * 4;
    bool staged = g_staged && (K == 64 || K == 128) && p.vec_ok && encode_tiled() != nullptr;
    p.nraw = tg::NRAW;
    if (staged && tg::SMEM_LIMIT - fixed - p.wbytes - p.nraw * chunk < 2 * (int)tg::STAGE) p.nraw = tg::NRAW - 1;
    if (staged && tg::SMEM_LIMIT - fixed - p.wbytes - p.nraw * chunk < 2 * (int)tg::STAGE) staged = false;
    const int raw_bytes = staged ? p.nraw * chunk : 0;
    int nst = (tg::SMEM_LIMIT - fixed - p.wbytes - raw_bytes) / (int)tg::STAGE;
    if (nst > tg::MAX_NST) nst = tg::MAX_NST;
    if (nst < 2) return fa_internal_fail(-1, "tg_linear: weights too large for the shared-memory ring");
    p.nst = nst;
    const size_t smem = (size_t)p.wbytes + (size_t)nst * tg::STAGE + fixed + raw_bytes;
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    CUtensorMap xmap;
    memset(&xmap, 0, sizeof(xmap));
    if (staged) {
        const cuuint64_t xdim[2] = {(cuuint64_t)K, (cuuint64_t)rows}, xstride[1] = {(cuuint64_t)ldx * 4};
        const cuuint32_t xbox[2] = {(cuuint32_t)K, (cuuint32_t)tg::RAW_ROWS}, estr[2] = {1, 1};
        if (!encode_tiled() ||
            encode_tiled()(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_x), xdim, xstride, xbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return fa_internal_fail(-2, "tg_linear: cuTensorMapEncodeTiled failed for the input rows");
    }
    if (staged && p.tma_out) tg::tg_linear_kernel<true, true><<<grid, tg::THREADS_STAGED, smem, (cudaStream_t)stream>>>(p, omap, rmap, xmap);
    else if (staged) tg::tg_linear_kernel<true, false><<<grid, tg::THREADS_STAGED, smem, (cudaStream_t)stream>>>(p, omap, rmap, xmap);
    else if (p.tma_out) tg::tg_linear_kernel<false, true><<<grid, tg::THREADS, smem, (cudaStream_t)stream>>>(p, omap, rmap, xmap);
    else tg::tg_linear_kernel<false, false><<<grid, tg::THREADS, smem, (cudaStream_t)stream>>>(p, omap, rmap, xmap);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_linear: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" size_t tg_wgrad_scratch_bytes(int a, int b) { return (size_t)192 * a * b * sizeof(float); }

extern "C" int tg_wgrad(const float *d_x, int ldx, int a, const float *d_y, int ldy, int b, long long rows, float *d_dw, int lddw,
                        int accumulate, void *d_scratch, uint32_t *d_status, void *stream) {
    if (!d_x || !d_y || !d_dw || !d_scratch || !d_status) return fa_internal_fail(-1, "tg_wgrad: NULL pointer");
    if (rows < 1 || a < 1 || a > 256 || b < 1 || b > 256 || ldx < a || ldy < b || lddw < b)
        return fa_internal_fail(-1, "tg_wgrad: need rows >= 1, 1 <= a, b <= 256, ld >= width");
    int sms = 0;
    if (int rc = prepare(&sms)) return rc;
    if (sms > 192) sms = 192;
    tg::WgParams p = {};
    p.x = d_x; p.y = d_y; p.partial = (float *)d_scratch; p.status = d_status; p.rows = rows; p.ldx = ldx; p.ldy = ldy;
    p.a = a; p.b = b; p.ab = (a + 127) / 128; p.bb = (b + 127) / 128;
    const int wr = (g_wg_rows == 32 || g_wg_rows == 64) ? g_wg_rows : (p.ab + p.bb >= 3 ? 32 : 64);
    p.n_steps = (int)((rows + wr - 1) / wr);
    p.vx = (ldx % 4 == 0) && (((uintptr_t)d_x & 15) == 0);
    p.vy = (ldy % 4 == 0) && (((uintptr_t)d_y & 15) == 0);
    p.lbo = g_wg_lbo; p.sbo = g_wg_sbo ? g_wg_sbo : (wr == 64 ? tg::WG<64, false>::LBO : tg::WG<32, false>::LBO);
    const int grid = p.n_steps < sms ? p.n_steps : sms;
    // staged form: both operands in whole 8-column chunks with 16-byte aligned rows
    CUtensorMap xmap, ymap;
    memset(&xmap, 0, sizeof(xmap));
    memset(&ymap, 0, sizeof(ymap));
    bool staged = g_wg_staged && p.vx && p.vy && a % 8 == 0 && b % 8 == 0 && encode_tiled() != nullptr;
    if (staged) {
        const cuuint32_t estr[2] = {1, 1};
        const cuuint64_t xdim[2] = {(cuuint64_t)a, (cuuint64_t)rows}, xstr[1] = {(cuuint64_t)ldx * 4};
        const cuuint64_t ydim[2] = {(cuuint64_t)b, (cuuint64_t)rows}, ystr[1] = {(cuuint64_t)ldy * 4};
        const cuuint32_t xbox[2] = {(cuuint32_t)(a < 128 ? a : 128), (cuuint32_t)wr}, ybox[2] = {(cuuint32_t)(b < 128 ? b : 128), (cuuint32_t)wr};
        staged = encode_tiled()(&xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_x), xdim, xstr, xbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS &&
                 encode_tiled()(&ymap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(d_y), ydim, ystr, ybox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
    }
    const cudaStream_t st = (cudaStream_t)stream;
    if (wr == 64 && staged) tg::tg_wgrad_kernel<64, true><<<grid, tg::THREADS_STAGED, tg::WG<64, true>::SMEM, st>>>(p, xmap, ymap);
    else if (wr == 64) tg::tg_wgrad_kernel<64, false><<<grid, tg::THREADS, tg::WG<64, false>::SMEM, st>>>(p, xmap, ymap);
    else if (staged) tg::tg_wgrad_kernel<32, true><<<grid, tg::THREADS_STAGED, tg::WG<32, true>::SMEM, st>>>(p, xmap, ymap);
    else tg::tg_wgrad_kernel<32, false><<<grid, tg::THREADS, tg::WG<32, false>::SMEM, st>>>(p, xmap, ymap);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_wgrad: launch: %s", cudaGetErrorString(e));
    tg::tg_reduce_kernel<<<(a * b * 8 + 255) / 256, 256, 0, (cudaStream_t)stream>>>((const float *)d_scratch, grid, a, b, d_dw, lddw, accumulate);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_wgrad: reduce launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int tg_adam_step(const TgTensor *tensors, int n_tensors, float lr, float beta1, float beta2, float eps, float max_norm,
                            const float *d_grad_scale, long long *d_step, float *d_scratch, float *d_total_norm, void *stream) {
    if (!tensors || !d_step || !d_scratch) return fa_internal_fail(-1, "tg_adam_step: NULL pointer");
    if (n_tensors < 1 || n_tensors > TG_MAX_TENSORS) return fa_internal_fail(-1, "tg_adam_step: 1 <= n_tensors <= %d", TG_MAX_TENSORS);
    tg::AdamParams p = {};
    for (int k = 0; k < n_tensors; ++k) {
        if (!tensors[k].p || !tensors[k].m || !tensors[k].v || tensors[k].numel < 0)
            return fa_internal_fail(-1, "tg_adam_step: tensor %d has a NULL parameter / moment pointer", k);
        p.t[k] = tensors[k];
    }
    p.n = n_tensors; p.lr = lr; p.b1 = beta1; p.b2 = beta2; p.eps = eps; p.max_norm = max_norm; p.grad_scale = d_grad_scale;
    p.step = d_step; p.scratch = d_scratch; p.total_out = d_total_norm;
    tg::tg_adam_norm_kernel<<<tg::ADAM_BLOCKS, 256, 0, (cudaStream_t)stream>>>(p);
    tg::tg_adam_update_kernel<<<tg::ADAM_BLOCKS, 256, 0, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_adam_step: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int tg_kernel_info(int which, int32_t *regs, int32_t *block, int32_t *smem) {
    if (int rc = prepare(nullptr)) return rc;
    cudaFuncAttributes at;
    const cudaError_t e = which == 0 ? cudaFuncGetAttributes(&at, tg::tg_linear_kernel<false, true>)
                        : (which == 2 ? cudaFuncGetAttributes(&at, tg::tg_linear_kernel<true, true>) : cudaFuncGetAttributes(&at, tg::tg_wgrad_kernel<64, true>));
    if (e != cudaSuccess) return fa_internal_fail(-2, "tg_kernel_info: %s", cudaGetErrorString(e));
    if (regs) *regs = at.numRegs;
    if (block) *block = which == 2 ? tg::THREADS_STAGED : tg::THREADS;
    if (smem) *smem = (which == 1 ? tg::WG<64, true>::SMEM : tg::SMEM_LIMIT) + (int)at.sharedSizeBytes;
    return 0;
}

extern "C" int tg_debug_wgrad_desc(uint32_t lbo, uint32_t sbo) {
    g_wg_lbo = lbo ? lbo : 128;
    g_wg_sbo = sbo;
    return 0;
}

extern "C" int tg_debug_wgrad_rows(int rows) {
    g_wg_rows = rows;
    return 0;
}

extern "C" int tg_debug_staged(int on) {
    g_staged = on != 0;
    return 0;
}

extern "C" int tg_debug_tma_out(int on) {
    g_tma_out = on != 0;
    return 0;
}

extern "C" int tg_debug_wgrad_staged(int on) {
    g_wg_staged = on != 0;
    return 0;
}
