// mp_policy.cu -- the reference's team policy forward (MPNN._fwd + act, mpnn.py:117-205) as ONE
// persistent sm_100a kernel: dense layers on tcgen05 tensor cores (fp16 operands, fp32 accumulators in
// tensor memory), everything else (input encoders, both attentions, heads, sampling) in fp32 on the
// CUDA cores of the same CTA.  C ABI: include/fortattack_policy.h.
//
// Algebra.  Both attentions are single-head and linear around their softmax, so the projections fold
// (done once on the host in float64, emergent-multiagent-strategies_b200/policy_kernel.py):
//   scores   q_a . k_b = (h_a Wq)(h_b Wk)^T = (h_a G) . h_b            G  = Wq Wk^T
//   message  (sum_b p_ab v_b) Wout U2^T = sum_b p_ab (h_b Wz)           Wz = Wv Wout U2^T
//   update   ReLU([h, msg] U^T + b) = ReLU(h U1^T + sum_b p_ab z_b + b) z = h Wz,  U = [U1 | U2]
// so one message round (mpnn.py:156-158) is ONE GEMM  [T | Z | Y] = h [G | Wz | U1^T]  (N = 384, K = 128)
// followed by a per-row epilogue; the 96 KB of round weights stay resident in shared memory for the whole
// kernel.  The opponent attention (mpnn.py:409-437) folds the same way into two 64 x 64 products.
//
// Tile = 128 rows = EPT environments x up to 5 agents, row r = a * EPT + e (EPT = 128 / max(n_own, n_opp)),
// so the rows an agent attends to (same e, other a) are in the same tile.  Thread r of the four
// "row" warps own row r for the whole network: it is tensor-memory lane r (tcgen05.ld 32x32b gives a
// thread its own accumulator row), they write row r of the next layer's fp16 A operand into shared
// memory, and they read other rows only for the two attentions.
//
// Two threads share a row: warp w and warp w+4 both map to tensor-memory lanes 32*(w%4)..+31 (the hardware's
// lane window of a warp is 32*(warp%4)), thread "half" h = w/4 handles columns [64h, 64h+64) of every
// per-row loop; attention scores are summed across the two halves through a small shared-memory exchange,
// and at the heads half 0 evaluates the value head while half 1 evaluates the action head and samples.
//
// Warp roles (320 threads, one CTA per SM, persistent over tiles):
//   warps 0-7  row threads (encoders, epilogues, attention, heads, sampling)
//   warp 8     weight producer: loads the resident round weights once, then streams the per-tile
//              weights (opponent attention, heads: 80 KB per tile) with cp.async.bulk into a ring
//   warp 9     MMA issuer: one thread issues every tcgen05.mma
// Synchronisation: a_ready (256 arrivals: "operand written / accumulator drained") row -> MMA,
// acc (tcgen05.commit) MMA -> row, named barrier 1 among the row threads.
//
// Shared memory: bufH [128 x 128 fp16] current h (A operand, also read by other rows for the scores),
// bufX [128 x 128 fp16] z rows (exchange between rows), 96 KB resident round weights, 10 KB fp32
// constants, NS x 16 KB ring.  Tensor memory: T | Z | Y = 384 of 512 columns.
//
// blob layout, fp16 part (byte offsets; B operands [N][K] in canonical K-major core-matrix order):
//        0 oppAttn: (Wkey Wquery^T)^T [64 x 64]      8192 oppAttn: (Wval Wout)^T [64 x 64]
//    16384 G^T [128 x 128]    49152 Wz^T [128 x 128]    81920 U1 [128 x 128]
//   114688 value_head.0.weight lo/hi [64 x 128] x2     147456 policy_head.0.weight lo/hi      (total 180224)
// fp32 part (float index): 0 encoder [64][8]={w0..w5,bias,0}  512 oppEncoder  1024 update.0.bias
//   1152 value_head.0.bias  1280 value_head.2.weight  1408 policy_head.0.bias  1536 dist.linear.weight^T [128][8]
//   2560 dist.linear.bias[8]  2568 value_head.2.bias
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/fortattack_policy.h"
#include "mp_umma.cuh"

int fa_internal_fail(int code, const char *fmt, ...);

namespace mp {

constexpr int STAGE_BYTES = 16384;             // one streamed weight chunk (N = 64 slice, K <= 128)
constexpr int NS = 3;                          // ring stages
constexpr uint32_t A_LBO = 128, A_SBO = 2048;  // [128 rows] x [K = 128] fp16 operands (activations, resident weights)
constexpr int NQ = 2;                          // threads per accumulator row (warps w, w+4, ... share a lane window);
                                               // measured: 4 is slower (68 vs 54 us at 3v3 x 16384): the phases are latency-bound
constexpr int ROW_WARPS = 4 * NQ, ROW_THREADS = 128 * NQ, THREADS = ROW_THREADS + 64, PRODUCER_WARP = ROW_WARPS,
              MMA_WARP = ROW_WARPS + 1;
constexpr int QC = 128 / NQ, QC64 = 64 / NQ;   // columns per thread of a 128-wide / 64-wide row segment
static_assert(NQ == 2 || NQ == 4, "two or four threads per row");
constexpr bool DW_IN_SMEM = NQ == 2;            // dist.linear.weight^T (4 KB): shared memory if it fits, else L1-cached global
constexpr int CONST_SMEM_FLOATS = DW_IN_SMEM ? MP_BLOB_CONST_FLOATS : 1552;
constexpr int OFF_H = 0, OFF_X = 32768, OFF_WRES = 65536, RES_BYTES = 98304, OFF_CONST = OFF_WRES + RES_BYTES,
              OFF_RING = OFF_CONST + (DW_IN_SMEM ? 10368 : 6272), SMEM_BYTES = OFF_RING + NS * STAGE_BYTES;
constexpr uint32_t BLOB_RES = 16384;           // blob offset of the resident part
constexpr int N_STREAM = 6;                    // streamed chunks per tile
constexpr int TMEM_COLS = 512;
constexpr uint32_t COL_T = 0, COL_Z = 128, COL_Y = 256;
static_assert(MP_BLOB_F16_BYTES == 180224 && CONST_SMEM_FLOATS * 4 <= OFF_RING - OFF_CONST, "blob layout");

// fp32 constant offsets
constexpr int C_ENC = 0, C_OENC = 512, C_UB = 1024, C_VB = 1152, C_VW = 1280, C_PB = 1408, C_DW = 1536, C_DB = 2560,
              C_VB2 = 2568;                      // blob indices
static_assert(C_VB2 == C_DB + 8 && C_VB2 < MP_BLOB_CONST_FLOATS, "value_head.2.bias follows dist.linear.bias");
constexpr int S_DB = DW_IN_SMEM ? C_DB : 1536, S_VB2 = S_DB + 8;   // shared-memory indices of dist.linear.bias, value_head.2.bias

struct Chunk {
    uint32_t off, bytes;   // in the blob
    uint32_t a_off;        // shared-memory byte offset of the A operand
    uint32_t b_sbo;        // SBO of the weight chunk (K * 16)
    uint16_t ksteps, tmem_col;
    uint8_t wait_a, commit_acc, pad0, pad1;
};
__constant__ Chunk c_tab[N_STREAM] = {
    {0u, 8192u, OFF_H, 1024u, 4, 0, 1, 1, 0, 0},          // T' = h0 (Wkey Wquery^T)          -> acc
    {8192u, 8192u, OFF_X, 1024u, 4, 64, 0, 1, 0, 0},      // z' = hOpp (Wval Wout)             -> acc
    {147456u, 16384u, OFF_H, 2048u, 8, 128, 1, 0, 0, 0},  // policy hidden lo / hi             -> acc
    {163840u, 16384u, OFF_H, 2048u, 8, 192, 0, 1, 0, 0},
    {114688u, 16384u, OFF_H, 2048u, 8, 0, 0, 0, 0, 0},    // value hidden lo / hi              -> acc
    {131072u, 16384u, OFF_H, 2048u, 8, 64, 0, 1, 0, 0}};

struct Params {
    const uint8_t *blob;
    const uint8_t *blobs[MP_MAX_ENSEMBLE];   // n_ckpt > 0: CTA c plays checkpoint c % n_ckpt with weights blobs[c % n_ckpt]
    int n_ckpt;
    const float *obs_own, *obs_opp;     // [n][E][6]
    const int64_t *action_in;
    float *value, *logp, *entropy, *logits;
    int64_t *action;
    int32_t *action_i32;
    uint32_t *status;
    uint64_t seed, offset, env_id0;
    const int32_t *env_sel;        // optional int32 [E]: outputs are written only for envs with env_sel[e] == sel_value
    const int32_t *env_order;      // optional int32 [E] + env_offsets int32 [K+1] (both on the device): this launch covers only the
    const int32_t *env_offsets;    //   envs env_order[env_offsets[sel_value] .. env_offsets[sel_value+1])  (compacted ensemble play)
    int32_t sel_value;
    unsigned long long *counter;   // optional device call counter {count, ticket}: overrides `offset`, bumped by the last CTA
    int n_own, n_opp, E, ept, n_tiles, mode;
    int opt;                       // bit 0: balanced heads, bit 1: next tile's opponent encoder under the heads' MMAs (MP_OPT, default 2)
#ifdef MP_TRACE                  // diagnostic build only (make trace): never in the product library
    unsigned long long *trace;   // optional: clock64() of row thread 0 of CTA 0 at every phase boundary of one of its tiles
    int trace_tile;              // which of CTA 0's tiles (0 = first)
#endif
};

__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        const uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ void bar_rows() { asm volatile("bar.sync 1, %0;" ::"n"(ROW_THREADS) : "memory"); }

__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t *>(&h);
}
__device__ __forceinline__ float2 unpack_h2(uint32_t u) {
    return __half22float2(*reinterpret_cast<const __half2 *>(&u));
}

// tensor-memory columns [taddr, taddr+NCOLS) of this thread's row -> fp16 row segment k0.. of a canonical operand
// (NCOLS = 16, 32 or 64: all loads are in flight before the one tcgen05.wait)
template <int NCOLS>
__device__ __forceinline__ void drain(uint32_t taddr, uint8_t *buf, int r, int k0) {
    uint32_t v[NCOLS / 16][16];
#pragma unroll
    for (int q = 0; q < NCOLS / 16; ++q) tmem_ld16(taddr + (uint32_t)(q * 16), v[q]);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < NCOLS / 16; ++q) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {          // 8 columns -> one 16-byte chunk
            uint32_t h[4];
#pragma unroll
            for (int j = 0; j < 4; ++j)
                h[j] = pack_h2(__uint_as_float(v[q][g * 8 + 2 * j]), __uint_as_float(v[q][g * 8 + 2 * j + 1]));
            *reinterpret_cast<uint4 *>(buf + canon_off(r, k0 + q * 16 + g * 8, A_LBO, A_SBO)) = make_uint4(h[0], h[1], h[2], h[3]);
        }
    }
}

// ReLU(W x + b) for the 6-float observation, outputs [j_lo, j_lo+QC64) -> fp16 row segment of buf (mpnn.py:127-128)
__device__ __forceinline__ void encode(const float *W8, const float (&o)[6], uint8_t *buf, int r, int j_lo) {
#pragma unroll
    for (int j0 = j_lo; j0 < j_lo + QC64; j0 += 8) {
        uint32_t h[4];
#pragma unroll
        for (int jj = 0; jj < 8; jj += 2) {
            float y[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const float4 w0 = *reinterpret_cast<const float4 *>(W8 + (j0 + jj + q) * 8);
                const float4 w1 = *reinterpret_cast<const float4 *>(W8 + (j0 + jj + q) * 8 + 4);
                float acc = w1.z;   // bias
                acc = fmaf(w0.x, o[0], acc); acc = fmaf(w0.y, o[1], acc); acc = fmaf(w0.z, o[2], acc);
                acc = fmaf(w0.w, o[3], acc); acc = fmaf(w1.x, o[4], acc); acc = fmaf(w1.y, o[5], acc);
                y[q] = fmaxf(acc, 0.0f);
            }
            h[jj >> 1] = pack_h2(y[0], y[1]);
        }
        *reinterpret_cast<uint4 *>(buf + canon_off(r, j0, A_LBO, A_SBO)) = make_uint4(h[0], h[1], h[2], h[3]);
    }
}

__device__ __forceinline__ void load_obs(const float *obs, int n, int a, int E, int eg, float (&o)[6]) {
    if (a < n && eg < E) {
        const float2 *p = reinterpret_cast<const float2 *>(obs + ((size_t)a * E + eg) * MP_OBS_DIM);
        const float2 v0 = p[0], v1 = p[1], v2 = p[2];
        o[0] = v0.x; o[1] = v0.y; o[2] = v1.x; o[3] = v1.y; o[4] = v2.x; o[5] = v2.y;
    } else {
#pragma unroll
        for (int i = 0; i < 6; ++i) o[i] = 0.0f;
    }
}

// s[b] += <16 fp32 accumulator columns of this row, fp16 segment (two 16-byte chunks from byte offset koff) of row b>
template <int CNT>
__device__ __forceinline__ void dot16(const uint32_t (&v)[16], const uint8_t *buf, const uint32_t (&rows)[MP_MAX_TEAM],
                                      uint32_t koff, float (&s)[MP_MAX_TEAM]) {
    uint4 u[CNT > 0 ? CNT : 1][2];
#pragma unroll
    for (int b = 0; b < CNT; ++b)
#pragma unroll
        for (int g = 0; g < 2; ++g) u[b][g] = *reinterpret_cast<const uint4 *>(buf + rows[b] + koff + (uint32_t)g * A_LBO);
#pragma unroll
    for (int b = 0; b < CNT; ++b) {
        // two scalar chains (even / odd columns); the packed FFMA2 form measured slower (56 vs 54 us at 3v3 x 16384)
        float acc0 = s[b], acc1 = 0.0f;
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const uint32_t w[4] = {u[b][g].x, u[b][g].y, u[b][g].z, u[b][g].w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_h2(w[j]);
                acc0 = fmaf(__uint_as_float(v[g * 8 + 2 * j]), f.x, acc0);
                acc1 = fmaf(__uint_as_float(v[g * 8 + 2 * j + 1]), f.y, acc1);
            }
        }
        s[b] = acc0 + acc1;
    }
}

// partial scores of this row against CNT other rows over NB*16 accumulator columns starting at taddr; the
// other rows' fp16 features start at byte offset koff0 (same columns)
template <int CNT, int NB>
__device__ __forceinline__ void scores(uint32_t taddr, const uint8_t *buf, const uint32_t (&rows)[MP_MAX_TEAM], uint32_t koff0,
                                       float (&s)[MP_MAX_TEAM]) {
    if constexpr (CNT > 0) {
    uint32_t v[NB][16];
#pragma unroll
    for (int c = 0; c < NB; ++c) tmem_ld16(taddr + (uint32_t)(c * 16), v[c]);
    tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < NB; ++c) dot16<CNT>(v[c], buf, rows, koff0 + (uint32_t)(c * 2) * A_LBO, s);
    }
}

// softmax over the first cnt entries (cnt = 0: a lone agent receives a zero message, mpnn.py:262-270)
__device__ __forceinline__ void softmax_small(float (&s)[MP_MAX_TEAM], int cnt, float scale) {
    float mx = -CUDART_INF_F;
#pragma unroll
    for (int b = 0; b < MP_MAX_TEAM; ++b) {
        s[b] = b < cnt ? s[b] * scale : -CUDART_INF_F;
        mx = fmaxf(mx, s[b]);
    }
    float sum = 0.0f;
#pragma unroll
    for (int b = 0; b < MP_MAX_TEAM; ++b) {
        s[b] = b < cnt ? __expf(s[b] - mx) : 0.0f;
        sum += s[b];
    }
    const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
#pragma unroll
    for (int b = 0; b < MP_MAX_TEAM; ++b) s[b] *= inv;
}

// acc[0..8) += sum_b p[b] * (8 fp16 values at byte offset koff of row b)
template <int CNT>
__device__ __forceinline__ void mix8(float (&acc)[8], const uint8_t *buf, const uint32_t (&rows)[MP_MAX_TEAM], uint32_t koff,
                                     const float (&p)[MP_MAX_TEAM]) {
    uint4 u[CNT > 0 ? CNT : 1];
#pragma unroll
    for (int b = 0; b < CNT; ++b) u[b] = *reinterpret_cast<const uint4 *>(buf + rows[b] + koff);
#pragma unroll
    for (int b = 0; b < CNT; ++b) {
        const uint32_t w[4] = {u[b].x, u[b].y, u[b].z, u[b].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float2 f = unpack_h2(w[j]);
            acc[2 * j] = fmaf(p[b], f.x, acc[2 * j]);
            acc[2 * j + 1] = fmaf(p[b], f.y, acc[2 * j + 1]);
        }
    }
}

// eOpp = sum_b p_b z'_b  (z' rows at bufX k = 64..127) -> this row of bufH, k = 64..127  (mpnn.py:142,432-437)
template <int CNT>
__device__ __forceinline__ void opp_message(const uint8_t *bufX, uint8_t *bufH, int r, const uint32_t (&rows)[MP_MAX_TEAM],
                                            const float (&p)[MP_MAX_TEAM], int kc_lo) {
#pragma unroll
    for (int kc = kc_lo; kc < kc_lo + 8 / NQ; ++kc) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        mix8<CNT>(acc, bufX, rows, (uint32_t)(8 + kc) * A_LBO, p);
        *reinterpret_cast<uint4 *>(bufH + canon_off(r, 64 + kc * 8, A_LBO, A_SBO)) =
            make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
    }
}

// h_new = ReLU(Y + sum_b p_b z_b + bias) for columns [c0, c0+QC) -> this row of bufH   (mpnn.py:157-158, folded)
template <int CNT>
__device__ __forceinline__ void update_row(uint32_t t_y, const uint8_t *bufX, uint8_t *bufH, int r, int c0,
                                           const uint32_t (&rows)[MP_MAX_TEAM], const float (&p)[MP_MAX_TEAM], const float *bias) {
    uint32_t y[QC / 16][16];
#pragma unroll
    for (int q = 0; q < QC / 16; ++q) tmem_ld16(t_y + (uint32_t)(c0 + q * 16), y[q]);
    tmem_ld_wait();
#pragma unroll
    for (int q = 0; q < QC / 16; ++q) {
#pragma unroll
        for (int g = 0; g < 2; ++g) {
            const int k = c0 + q * 16 + g * 8;
            const float4 b0 = *reinterpret_cast<const float4 *>(bias + k), b1 = *reinterpret_cast<const float4 *>(bias + k + 4);
            float acc[8] = {__uint_as_float(y[q][g * 8]) + b0.x,     __uint_as_float(y[q][g * 8 + 1]) + b0.y,
                            __uint_as_float(y[q][g * 8 + 2]) + b0.z, __uint_as_float(y[q][g * 8 + 3]) + b0.w,
                            __uint_as_float(y[q][g * 8 + 4]) + b1.x, __uint_as_float(y[q][g * 8 + 5]) + b1.y,
                            __uint_as_float(y[q][g * 8 + 6]) + b1.z, __uint_as_float(y[q][g * 8 + 7]) + b1.w};
            mix8<CNT>(acc, bufX, rows, (uint32_t)(k >> 3) * A_LBO, p);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = fmaxf(acc[j], 0.0f);
            *reinterpret_cast<uint4 *>(bufH + canon_off(r, k, A_LBO, A_SBO)) =
                make_uint4(pack_h2(acc[0], acc[1]), pack_h2(acc[2], acc[3]), pack_h2(acc[4], acc[5]), pack_h2(acc[6], acc[7]));
        }
    }
}

// the NQ threads of a row add their partial scores: write -> barrier -> read the partners' (the barrier also
// publishes whatever the row threads wrote to shared memory before it)
__device__ __forceinline__ void combine_scores(float (&s)[MP_MAX_TEAM], float (*sc)[128][MP_MAX_TEAM], int q, int r) {
#pragma unroll
    for (int b = 0; b < MP_MAX_TEAM; ++b) sc[q][r][b] = s[b];
    bar_rows();
#pragma unroll
    for (int b = 0; b < MP_MAX_TEAM; ++b) {
        float t = 0.0f;
#pragma unroll
        for (int k = 0; k < NQ; ++k) t += sc[k][r][b];     // same order in every thread of the row: identical sums
        s[b] = t;
    }
}

// run CALL(CNT) with the team size as a compile-time constant (branch-free, fully unrolled inner loops)
#define MP_DISPATCH(cnt, CALL)            \
    switch (cnt) {                        \
        case 0: { CALL(0); } break;       \
        case 1: { CALL(1); } break;       \
        case 2: { CALL(2); } break;       \
        case 3: { CALL(3); } break;       \
        case 4: { CALL(4); } break;       \
        default: { CALL(5); } break;      \
    }

__global__ void __launch_bounds__(THREADS, 1) mp_policy_kernel(const Params p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // bar_acc[2]: accumulator hand-offs alternate between two barriers (commit k -> bar_acc[k & 1], phase (k >> 1) & 1), so the
    // MMA issuer may finish TWO groups before the rows look at the first (policy and value hidden layers while the rows encode
    // the next tile's opponents): with one barrier the second completion flips the phase parity back and the rows' wait for
    // the first never returns
    __shared__ __align__(8) uint64_t bar_full[NS], bar_empty[NS], bar_a_ready, bar_acc[2], bar_res;
    __shared__ uint32_t tmem_base_s;
    __shared__ float sc[NQ][128][MP_MAX_TEAM];         // partial sums exchanged between the NQ threads of a row
    uint8_t *bufH = smem + OFF_H, *bufX = smem + OFF_X;
    float *C = reinterpret_cast<float *>(smem + OFF_CONST);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // ensemble launch: CTA c serves checkpoint c % n_ckpt (its weights stay resident for the whole launch) and strides
    // over that checkpoint's tiles together with the other CTAs of the same checkpoint
    const int ckpt = p.n_ckpt > 0 ? (int)(blockIdx.x % (unsigned)p.n_ckpt) : p.sel_value;
    const uint8_t *blob = p.n_ckpt > 0 ? p.blobs[ckpt] : p.blob;
    const int tile0 = p.n_ckpt > 0 ? (int)(blockIdx.x / (unsigned)p.n_ckpt) : (int)blockIdx.x;
    const int tile_stride = p.n_ckpt > 0 ? ((int)gridDim.x - ckpt + p.n_ckpt - 1) / p.n_ckpt : (int)gridDim.x;

    if (tid == 0) {
        for (int s = 0; s < NS; ++s) { mbar_init(&bar_full[s], 1); mbar_init(&bar_empty[s], 1); }
        mbar_init(&bar_a_ready, ROW_THREADS);
        mbar_init(&bar_acc[0], 1);
        mbar_init(&bar_acc[1], 1);
        mbar_init(&bar_res, 1);
        mbar_fence_init();
    }
    if (warp == PRODUCER_WARP) tmem_alloc<TMEM_COLS>(&tmem_base_s);
    {   // fp32 constants: plain loads, once per CTA
        const float4 *src = reinterpret_cast<const float4 *>(blob + MP_BLOB_F16_BYTES);
        float4 *dst = reinterpret_cast<float4 *>(C);
        if (DW_IN_SMEM) {
            for (int i = tid; i < MP_BLOB_CONST_FLOATS / 4; i += THREADS) dst[i] = src[i];
        } else {
            for (int i = tid; i < C_DW / 4; i += THREADS) dst[i] = src[i];                   // encoders, biases, value_head.2
            if (tid < 4) dst[S_DB / 4 + tid] = src[C_DB / 4 + tid];                           // dist.linear.bias, value_head.2.bias
        }
    }
    // call counter of the sampling stream: read before any CTA can have finished (the bump below happens after ALL
    // CTAs are done), so launches replayed from a CUDA graph still draw fresh numbers
    const uint64_t call_offset = p.counter != nullptr ? (uint64_t)p.counter[0] : p.offset;
    // compacted launch: the list of environments (and with it the tile count) is read from device memory, so the host
    // never has to know how many environments each checkpoint currently plays (no synchronisation, graph-replayable)
    const int32_t *env_list = nullptr;
    int n_list = p.E, n_tiles = p.n_tiles;
    if (p.env_order != nullptr) {
        const int lo = p.env_offsets[ckpt];
        n_list = p.env_offsets[ckpt + 1] - lo;
        env_list = p.env_order + lo;
        n_tiles = (n_list + p.ept - 1) / p.ept;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base_s;

    if (warp == PRODUCER_WARP) {
        // ================= weight producer =================
        if (lane == 0) {
            mbar_expect_tx(&bar_res, RES_BYTES);           // round weights: resident for the whole kernel
            for (int i = 0; i < RES_BYTES / STAGE_BYTES; ++i)
                bulk_g2s(smem + OFF_WRES + i * STAGE_BYTES, blob + BLOB_RES + i * STAGE_BYTES, STAGE_BYTES, &bar_res);
            uint32_t g = 0;
            for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
                for (int c = 0; c < N_STREAM; ++c, ++g) {
                    const uint32_t s = g % NS, ph = (g / NS) & 1u;
                    mbar_wait(&bar_empty[s], ph ^ 1u, p.status, 3);
#ifdef MP_TRACE
                    if (p.trace != nullptr && blockIdx.x == 0 && tile == 0) p.trace[64 + c] = clock64();
#endif
                    mbar_expect_tx(&bar_full[s], c_tab[c].bytes);
                    bulk_g2s(smem + OFF_RING + s * STAGE_BYTES, blob + c_tab[c].off, c_tab[c].bytes, &bar_full[s]);
                }
            }
        }
        __syncwarp();
    } else if (warp == MMA_WARP) {
        // ================= MMA issuer =================
        // The whole warp runs this code converged (waits included) and one elected lane issues the tcgen05
        // instructions: everything around them is warp-uniform, so descriptors live in uniform registers and the
        // issue rate keeps up with the tensor pipe.
        const uint32_t idesc64 = idesc_f16(128, 64), idesc128 = idesc_f16(128, 128);
        const uint32_t sbase = smem_u32(smem);
        const uint64_t hdesc = smem_desc(sbase + OFF_H, A_LBO, A_SBO), wdesc = smem_desc(sbase + OFF_WRES, A_LBO, A_SBO);
        uint32_t g = 0, pa = 0, ac = 0;                    // ac: accumulator groups committed so far (warp-uniform)
        auto stream_chunk = [&](int c) {
            const Chunk ch = c_tab[c];
            const uint32_t s = g % NS, ph = (g / NS) & 1u;
            if (ch.wait_a) { mbar_wait(&bar_a_ready, pa, p.status, 4); pa ^= 1u; }
            mbar_wait(&bar_full[s], ph, p.status, 5);
            tc_fence_after();
            // descriptors differ only in the start-address field (+256 bytes = +16 per K step)
            const uint64_t ad0 = smem_desc(sbase + ch.a_off, A_LBO, A_SBO);
            const uint64_t bd0 = smem_desc(sbase + OFF_RING + s * STAGE_BYTES, A_LBO, ch.b_sbo);
            const uint32_t dcol = tmem + ch.tmem_col;
            if (elect_one()) {
                if (ch.ksteps == 8) {
#pragma unroll
                    for (uint32_t k = 0; k < 8; ++k) umma_f16(dcol, ad0 + k * 16u, bd0 + k * 16u, idesc64, k > 0);
                } else {
#pragma unroll
                    for (uint32_t k = 0; k < 4; ++k) umma_f16(dcol, ad0 + k * 16u, bd0 + k * 16u, idesc64, k > 0);
                }
                umma_commit(&bar_empty[s]);
                if (ch.commit_acc) umma_commit(&bar_acc[ac & 1u]);
#ifdef MP_TRACE
                if (p.trace != nullptr && blockIdx.x == 0 && g < 6) p.trace[72 + c] = clock64();
#endif
            }
            __syncwarp();
            ac += ch.commit_acc;
            ++g;
        };
        mbar_wait(&bar_res, 0, p.status, 2);
        for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
            stream_chunk(0);
            stream_chunk(1);
            for (int round = 0; round < 3; ++round) {      // [T | Z | Y] = h [G | Wz | U1^T]
                mbar_wait(&bar_a_ready, pa, p.status, 4);
                pa ^= 1u;
                tc_fence_after();
                if (elect_one()) {
#pragma unroll
                    for (uint32_t j = 0; j < 3; ++j) {
#pragma unroll
                        for (uint32_t k = 0; k < 8; ++k)
                            umma_f16(tmem + j * 128u, hdesc + k * 16u, wdesc + (j * 32768u >> 4) + k * 16u, idesc128, k > 0);
                        if (j != 1) umma_commit(&bar_acc[(ac + (j >> 1)) & 1u]);   // T ready (the rows start the scores), then Z | Y
                    }
                }
                __syncwarp();
                ac += 2;
            }
            for (int c = 2; c < N_STREAM; ++c) stream_chunk(c);
        }
    } else {
        // ================= row threads: row r, column slice q of NQ =================
        const int q = warp >> 2, r = (warp & 3) * 32 + lane;
        const int c0 = q * QC;                                     // this thread's columns of a 128-wide row
        const uint32_t trow = tmem + ((uint32_t)((warp & 3) * 32) << 16);
        const int ept = p.ept, n_own = p.n_own, n_opp = p.n_opp;
        const int a = r / ept, e = r - a * ept;
        // byte offsets of the rows this row attends to: opponents (b, e) for every b; team mates (b, e), b != a
        uint32_t rowopp[MP_MAX_TEAM], rowoth[MP_MAX_TEAM];
#pragma unroll
        for (int b = 0; b < MP_MAX_TEAM; ++b) {
            const int rb = min(b * ept + e, 127);
            rowopp[b] = (uint32_t)(rb >> 3) * A_SBO + (uint32_t)(rb & 7) * 16u;
            const int ro = min((b + (b >= a ? 1 : 0)) * ept + e, 127);
            rowoth[b] = (uint32_t)(ro >> 3) * A_SBO + (uint32_t)(ro & 7) * 16u;
        }
        const int n_oth = n_own - 1;
        const float *DW = DW_IN_SMEM ? C + C_DW : reinterpret_cast<const float *>(blob + MP_BLOB_F16_BYTES) + C_DW;   // [128][8]
        uint32_t pc = 0;
#ifdef MP_TRACE
        int ti = 0;
#endif
#define ARRIVE_A() do { tc_fence_before(); fence_async_smem(); mbar_arrive(&bar_a_ready); } while (0)
#ifdef MP_TRACE
#define TS() do { if (p.trace != nullptr && tid == 0 && blockIdx.x == 0 && tile == tile0 + p.trace_tile * tile_stride && ti < 96) p.trace[ti++] = clock64(); } while (0)
#else
#define TS() do { } while (0)
#endif
#define WAIT_ACC(code) do { mbar_wait(&bar_acc[pc & 1u], (pc >> 1) & 1u, p.status, code); ++pc; tc_fence_after(); } while (0)

        float o_own[6], o_opp[6];                                  // this tile's observations (prefetched one tile ahead)
        // global env id of list entry li (p.E = "none": load_obs and the output guard treat it as out of range)
        auto env_of = [&](int li) { return li < n_list ? (env_list != nullptr ? env_list[li] : li) : p.E; };
        load_obs(p.obs_own, n_own, a, p.E, env_of(tile0 * ept + e), o_own);
        load_obs(p.obs_opp, n_opp, a, p.E, env_of(tile0 * ept + e), o_opp);
        for (int tile = tile0; tile < n_tiles; tile += tile_stride) {
            const int eg = env_of(tile * ept + e);
            TS();
            // ---- input encoders (mpnn.py:127-128): h0 -> bufH[:, 0:64), hOpp -> bufX[:, 0:64); QC64 outputs per thread
            encode(C + C_ENC, o_own, bufH, r, q * QC64);
            if (tile == tile0 || !(p.opt & 2)) encode(C + C_OENC, o_opp, bufX, r, q * QC64);   // later tiles: encoded under the previous tile's heads
            ARRIVE_A();
            TS();
            // ---- attention over the opponents (mpnn.py:409-437), folded: T' = h0 G', z' = hOpp Wz' -------
            {
                float s[MP_MAX_TEAM] = {0.f, 0.f, 0.f, 0.f, 0.f};
                WAIT_ACC(6);                                      // T'
                TS();
#define CALL(N) scores<N, QC64 / 16>(trow + COL_T + q * QC64, bufX, rowopp, (uint32_t)(q * QC64 >> 3) * A_LBO, s)
                MP_DISPATCH(n_opp, CALL)                          // T' . hOpp_b over this thread's QC64 features
#undef CALL
                WAIT_ACC(7);                                      // z'
                drain<QC64>(trow + 64 + q * QC64, bufX, r, 64 + q * QC64);   // z' of opponent row r -> bufX[:, 64:128)
                combine_scores(s, sc, q, r);                      // (barrier: every z' row is in bufX)
                softmax_small(s, n_opp, 0.125f);                  // 1/sqrt(64)
                TS();
#define CALL(N) opp_message<N>(bufX, bufH, r, rowopp, s, q * (8 / NQ))
                MP_DISPATCH(n_opp, CALL)                          // h = [h0 | eOpp]
#undef CALL
            }
            ARRIVE_A();
            TS();

            // ---- three message-passing rounds (mpnn.py:156-158) ------------------------------------------
            for (int round = 0; round < 3; ++round) {
                if (round == 2) {      // the next tile's observations: the loads fly during the last round and the heads
                    const int nt = tile + tile_stride;
                    if (nt < n_tiles) {
                        load_obs(p.obs_own, n_own, a, p.E, env_of(nt * ept + e), o_own);
                        load_obs(p.obs_opp, n_opp, a, p.E, env_of(nt * ept + e), o_opp);
                    }
                }
                float s[MP_MAX_TEAM] = {0.f, 0.f, 0.f, 0.f, 0.f};
                WAIT_ACC(8);                                      // T ready; Z | Y are still being computed
                TS();
#define CALL(N) scores<N, QC / 16>(trow + COL_T + c0, bufH, rowoth, (uint32_t)(c0 >> 3) * A_LBO, s)
                MP_DISPATCH(n_oth, CALL)                          // (h_a G) . h_b against the team mates' h rows
#undef CALL
                TS();
                WAIT_ACC(9);                                      // Z | Y ready
                drain<QC>(trow + COL_Z + c0, bufX, r, c0);       // this slice of z -> bufX
                combine_scores(s, sc, q, r);                      // (barrier: z rows visible; nobody reads bufH any more)
                softmax_small(s, n_oth, 0.08838834764831845f);    // 1/sqrt(128); no self message (mpnn.py:297-298)
                TS();
#define CALL(N) update_row<N>(trow + COL_Y, bufX, bufH, r, c0, rowoth, s, C + C_UB)
                MP_DISPATCH(n_oth, CALL)
#undef CALL
                ARRIVE_A();
                TS();
            }
            // this tile's identity, before the observation registers move on (eg / a are per-tile constants already)
            if ((p.opt & 2) && tile + tile_stride < n_tiles) {
                // The heads' MMAs read bufH only; bufX is free once every row has finished the last update (which reads the z
                // rows of its team mates): one barrier, then the opponents' encoder of the NEXT tile runs while the tensor pipe
                // works on the heads (ncu / phase trace: the rows idled ~1.8k cycles here).
                bar_rows();
                encode(C + C_OENC, o_opp, bufX, r, q * QC64);
            }

            // ---- heads (mpnn.py:174-205).  Threads q < NQ/2 evaluate the value head, the others policy_head +
            // dist.linear, each over a slice of the 128 hidden units; partial sums meet in `sc`, thread NQ/2 of the row
            // then finishes the Categorical.  The MMA issuer commits the policy hidden layer first, then the value
            // hidden layer: the logits are computed while the value weights are still streaming in.
            constexpr int HG = NQ / 2, HC = 128 / HG;             // threads per head, hidden units per thread
            WAIT_ACC(11);
            TS();
            const bool live = a < n_own && eg < p.E && (p.env_sel == nullptr || p.env_sel[eg] == p.sel_value);
            const size_t row = (size_t)a * p.E + eg;
            float value = 0.0f, lg[MP_ACTIONS] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (NQ == 2 && (p.opt & 1)) {
                // balanced form: each of the two threads of a row takes 64 hidden units of the action head AND of the value head
                // (before: one thread 128 x 8 FMAs, the other 128 FMAs -- the row waited for the long one), partial sums meet in sc
                const int k0 = q * 64;
#pragma unroll 1
                for (int c = 0; c < 64; c += 16) {
                    uint32_t w[16];
                    tmem_ld16(trow + (uint32_t)(128 + k0 + c), w);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = k0 + c + j;
                        const float x = fmaxf(__uint_as_float(w[j]) + C[C_PB + k], 0.0f);
                        const float4 d0 = *reinterpret_cast<const float4 *>(DW + k * 8);
                        const float4 d1 = *reinterpret_cast<const float4 *>(DW + k * 8 + 4);
                        lg[0] = fmaf(x, d0.x, lg[0]); lg[1] = fmaf(x, d0.y, lg[1]); lg[2] = fmaf(x, d0.z, lg[2]);
                        lg[3] = fmaf(x, d0.w, lg[3]); lg[4] = fmaf(x, d1.x, lg[4]); lg[5] = fmaf(x, d1.y, lg[5]);
                        lg[6] = fmaf(x, d1.z, lg[6]); lg[7] = fmaf(x, d1.w, lg[7]);
                    }
                }
                WAIT_ACC(12);    // value hidden layer; also: it still reads bufH, nobody may start the next tile before it is done
#pragma unroll 1
                for (int c = 0; c < 64; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(trow + (uint32_t)(k0 + c), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = k0 + c + j;
                        value = fmaf(fmaxf(__uint_as_float(v[j]) + C[C_VB + k], 0.0f), C[C_VW + k], value);
                    }
                }
                if (q == 0) {                                     // hand the partial logits to the thread that finishes the Categorical
#pragma unroll
                    for (int k = 0; k < MP_ACTIONS; ++k) sc[k / MP_MAX_TEAM][r][k % MP_MAX_TEAM] = lg[k];
                } else {
                    sc[1][r][MP_MAX_TEAM - 1] = value;            // (logits use sc[0][r][0..4], sc[1][r][0..2])
                }
                bar_rows();
                if (q == 0) {
                    value += sc[1][r][MP_MAX_TEAM - 1];
                } else {
#pragma unroll
                    for (int k = 0; k < MP_ACTIONS; ++k) lg[k] = sc[k / MP_MAX_TEAM][r][k % MP_MAX_TEAM] + lg[k];
                }
            } else {
            if (q >= HG) {
                const int k0 = (q - HG) * HC;
#pragma unroll 1
                for (int c = 0; c < HC; c += 16) {
                    uint32_t w[16];
                    tmem_ld16(trow + (uint32_t)(128 + k0 + c), w);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = k0 + c + j;
                        const float x = fmaxf(__uint_as_float(w[j]) + C[C_PB + k], 0.0f);
                        const float4 d0 = *reinterpret_cast<const float4 *>(DW + k * 8);
                        const float4 d1 = *reinterpret_cast<const float4 *>(DW + k * 8 + 4);
                        lg[0] = fmaf(x, d0.x, lg[0]); lg[1] = fmaf(x, d0.y, lg[1]); lg[2] = fmaf(x, d0.z, lg[2]);
                        lg[3] = fmaf(x, d0.w, lg[3]); lg[4] = fmaf(x, d1.x, lg[4]); lg[5] = fmaf(x, d1.y, lg[5]);
                        lg[6] = fmaf(x, d1.z, lg[6]); lg[7] = fmaf(x, d1.w, lg[7]);
                    }
                }
                if (q > HG) {                                     // (NQ = 4) second policy thread: hand the partial logits over
#pragma unroll
                    for (int k = 0; k < MP_ACTIONS; ++k) sc[2 + k / MP_MAX_TEAM][r][k % MP_MAX_TEAM] = lg[k];
                }
            }
            WAIT_ACC(12);    // value hidden layer; also: it still reads bufH, nobody may start the next tile before it is done
            if (q < HG) {
                const int k0 = q * HC;
#pragma unroll 1
                for (int c = 0; c < HC; c += 16) {
                    uint32_t v[16];
                    tmem_ld16(trow + (uint32_t)(k0 + c), v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        const int k = k0 + c + j;
                        value = fmaf(fmaxf(__uint_as_float(v[j]) + C[C_VB + k], 0.0f), C[C_VW + k], value);
                    }
                }
                if (q > 0) sc[1][r][0] = value;
            }
            if (NQ > 2) bar_rows();
            }
            if (q == 0) {
                if (NQ > 2) value += sc[1][r][0];
                if (live && p.value) p.value[row] = value + C[S_VB2];
            } else if (q == HG && live) {
#pragma unroll
                for (int k = 0; k < MP_ACTIONS; ++k) {
                    if (NQ > 2) lg[k] += sc[2 + k / MP_MAX_TEAM][r][k % MP_MAX_TEAM];
                    lg[k] += C[S_DB + k];
                }
                float mx = lg[0];
#pragma unroll
                for (int k = 1; k < MP_ACTIONS; ++k) mx = fmaxf(mx, lg[k]);
                float pr[MP_ACTIONS], sum = 0.0f;
#pragma unroll
                for (int k = 0; k < MP_ACTIONS; ++k) { pr[k] = expf(lg[k] - mx); sum += pr[k]; }
                const float lse = mx + logf(sum), inv = 1.0f / sum;
                int act = 0;
                if (p.mode == MP_MODE_SAMPLE) {
                    const uint64_t env = p.env_id0 + (uint64_t)eg;
                    uint32_t ctr[4] = {(uint32_t)env, (uint32_t)(env >> 32), (uint32_t)call_offset,
                                       (uint32_t)(call_offset >> 32) ^ ((uint32_t)a << 24)};
                    philox4x32_10((uint32_t)p.seed, (uint32_t)(p.seed >> 32), ctr);
                    const float u = ((float)(ctr[0] >> 8) + 0.5f) * (1.0f / 16777216.0f);
                    // inverse CDF over softmax(logits): the distribution of dist.sample() (distributions.py:11-13)
                    float cum = 0.0f;
                    bool found = false;
                    act = MP_ACTIONS - 1;
#pragma unroll
                    for (int k = 0; k < MP_ACTIONS - 1; ++k) {
                        cum += pr[k] * inv;
                        if (!found && u < cum) { act = k; found = true; }
                    }
                } else if (p.mode == MP_MODE_ARGMAX) {
#pragma unroll
                    for (int k = 1; k < MP_ACTIONS; ++k) if (lg[k] > lg[act]) act = k;
                } else {
                    act = (int)p.action_in[row];
                    act = act < 0 ? 0 : (act > MP_ACTIONS - 1 ? MP_ACTIONS - 1 : act);
                }
                float lp = 0.0f, ent = 0.0f;
#pragma unroll
                for (int k = 0; k < MP_ACTIONS; ++k) {
                    const float l = lg[k] - lse;
                    if (k == act) lp = l;
                    ent -= pr[k] * inv * l;
                }
                if (p.action) p.action[row] = act;
                if (p.action_i32) p.action_i32[row] = act;
                if (p.logp) p.logp[row] = lp;
                if (p.entropy) p.entropy[row] = ent;
                if (p.logits) {
                    float4 *o = reinterpret_cast<float4 *>(p.logits + row * MP_ACTIONS);
                    o[0] = make_float4(lg[0], lg[1], lg[2], lg[3]);
                    o[1] = make_float4(lg[4], lg[5], lg[6], lg[7]);
                }
            }
            TS();
        }
#undef TS
#undef ARRIVE_A
#undef WAIT_ACC
    }
    tc_fence_before();
    __syncthreads();
    if (warp == PRODUCER_WARP) tmem_dealloc<TMEM_COLS>(tmem);
    if (p.counter != nullptr && tid == 0) {           // the last CTA to finish advances the call counter
        __threadfence();
        if (atomicAdd(&p.counter[1], 1ull) == (unsigned long long)gridDim.x - 1ull) {
            p.counter[1] = 0ull;
            p.counter[0] = call_offset + 1ull;
        }
    }
}

}  // namespace mp

// ---- host side ------------------------------------------------------------------------------------
namespace {
// per-device one-time setup (cudaFuncSetAttribute is a per-device setting; the SM count sizes the persistent grid)
constexpr int MAX_DEVICES = 64;
bool g_attr_done[MAX_DEVICES] = {};
int g_sms[MAX_DEVICES] = {};
#ifdef MP_TRACE
unsigned long long *g_trace = nullptr;
int g_trace_tile = 0;
#endif

// MP_OPT=0..3 switches two measures (measured, profiles/r3f_policy_time.log, 3v3 x 16384 / x 65536 envs, us per forward):
//   bit 0  balanced heads (each thread of a row takes 64 hidden units of both heads)      55.6 -> 56.2 / 178.3 -> 181.2  (off)
//   bit 1  the next tile's opponent encoder runs while the tensor pipe works on the heads 55.6 -> 54.9 / 178.3 -> 174.6  (on)
int g_opt() {
    static int v = -1;
    if (v < 0) {
        const char *ev = getenv("MP_OPT");
        v = ev ? (atoi(ev) & 3) : 2;
    }
    return v;
}

int prepare(int *sms) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess || dev < 0 || dev >= MAX_DEVICES) return fa_internal_fail(-3, "mp_policy_kernel: no usable CUDA device");
    if (!g_attr_done[dev]) {
        e = cudaFuncSetAttribute(mp::mp_policy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, mp::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaDeviceGetAttribute(&g_sms[dev], cudaDevAttrMultiProcessorCount, dev);
        if (e != cudaSuccess) return fa_internal_fail(-2, "mp_policy_kernel: device setup: %s", cudaGetErrorString(e));
        g_attr_done[dev] = true;
    }
    if (sms) *sms = g_sms[dev];
    return 0;
}
}  // namespace

extern "C" int mp_forward(const void *d_blob, const float *d_obs_own, const float *d_obs_opp, int n_own, int n_opp,
                          int n_envs, int mode, uint64_t seed, uint64_t offset, uint64_t *d_counter, uint64_t env_id0,
                          const int64_t *d_action_in,
                          float *d_value, int64_t *d_action, int32_t *d_action_i32, float *d_logp, float *d_entropy,
                          float *d_logits, const int32_t *d_env_sel, int32_t sel_value, const int32_t *d_env_order,
                          const int32_t *d_env_offsets, uint32_t *d_status, void *stream) {
    if (!d_blob || !d_obs_own || !d_obs_opp || !d_status) return fa_internal_fail(-1, "mp_forward: NULL pointer");
    if (n_own < 1 || n_own > MP_MAX_TEAM || n_opp < 1 || n_opp > MP_MAX_TEAM || n_envs < 1)
        return fa_internal_fail(-1, "mp_forward: team sizes must be 1..%d and n_envs >= 1", MP_MAX_TEAM);
    if (mode < 0 || mode > 2 || (mode == MP_MODE_EVAL && !d_action_in)) return fa_internal_fail(-1, "mp_forward: bad mode");
    if (((uintptr_t)d_blob & 15) || ((uintptr_t)d_obs_own & 7) || ((uintptr_t)d_obs_opp & 7) || ((uintptr_t)d_logits & 15))
        return fa_internal_fail(-4, "mp_forward: blob/logits must be 16-byte, observations 8-byte aligned");
    int sms = 0;
    if (int rc = prepare(&sms)) return rc;
    mp::Params p = {};
    p.blob = (const uint8_t *)d_blob; p.obs_own = d_obs_own; p.obs_opp = d_obs_opp; p.action_in = d_action_in;
    p.value = d_value; p.logp = d_logp; p.entropy = d_entropy; p.logits = d_logits; p.action = d_action;
    p.action_i32 = d_action_i32; p.status = d_status; p.seed = seed; p.offset = offset; p.env_id0 = env_id0; p.counter = (unsigned long long *)d_counter; p.env_sel = d_env_sel; p.sel_value = sel_value;
    p.env_order = d_env_order; p.env_offsets = d_env_offsets;
    if ((d_env_order == nullptr) != (d_env_offsets == nullptr) || (d_env_order != nullptr && sel_value < 0))
        return fa_internal_fail(-1, "mp_forward: d_env_order and d_env_offsets come together, with sel_value >= 0");
    p.n_own = n_own; p.n_opp = n_opp; p.E = n_envs; p.mode = mode; p.opt = g_opt();
#ifdef MP_TRACE
    p.trace = g_trace; p.trace_tile = g_trace_tile;
#endif
    p.ept = 128 / (n_own > n_opp ? n_own : n_opp);
    p.n_tiles = (n_envs + p.ept - 1) / p.ept;
    const int grid = p.n_tiles < sms ? p.n_tiles : sms;
    mp::mp_policy_kernel<<<grid, mp::THREADS, mp::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "mp_forward: launch: %s", cudaGetErrorString(e));
    return 0;
}

// One launch for an ensemble of K frozen checkpoints: CTA c serves checkpoint c % K over the environments
// d_env_order[d_env_offsets[k] .. d_env_offsets[k+1]).
extern "C" int mp_forward_ensemble(const void *const *d_blobs, int n_ckpt, const float *d_obs_own, const float *d_obs_opp,
                                   int n_own, int n_opp, int n_envs, int mode, uint64_t seed, uint64_t offset,
                                   uint64_t *d_counter, uint64_t env_id0, float *d_value, int64_t *d_action,
                                   int32_t *d_action_i32, float *d_logp, const int32_t *d_env_order,
                                   const int32_t *d_env_offsets, uint32_t *d_status, void *stream) {
    if (!d_blobs || !d_obs_own || !d_obs_opp || !d_status || !d_env_order || !d_env_offsets)
        return fa_internal_fail(-1, "mp_forward_ensemble: NULL pointer");
    if (n_ckpt < 1 || n_ckpt > MP_MAX_ENSEMBLE) return fa_internal_fail(-1, "mp_forward_ensemble: 1 <= n_ckpt <= %d", MP_MAX_ENSEMBLE);
    if (n_own < 1 || n_own > MP_MAX_TEAM || n_opp < 1 || n_opp > MP_MAX_TEAM || n_envs < 1 || (mode != MP_MODE_SAMPLE && mode != MP_MODE_ARGMAX))
        return fa_internal_fail(-1, "mp_forward_ensemble: bad team sizes, n_envs or mode");
    int sms = 0;
    if (int rc = prepare(&sms)) return rc;
    mp::Params p = {};
    for (int k = 0; k < n_ckpt; ++k) {
        if (!d_blobs[k] || ((uintptr_t)d_blobs[k] & 15)) return fa_internal_fail(-4, "mp_forward_ensemble: blob %d NULL or not 16-byte aligned", k);
        p.blobs[k] = (const uint8_t *)d_blobs[k];
    }
    p.n_ckpt = n_ckpt; p.blob = p.blobs[0];
    p.obs_own = d_obs_own; p.obs_opp = d_obs_opp; p.value = d_value; p.logp = d_logp; p.action = d_action;
    p.action_i32 = d_action_i32; p.status = d_status; p.seed = seed; p.offset = offset; p.env_id0 = env_id0;
    p.counter = (unsigned long long *)d_counter; p.env_order = d_env_order; p.env_offsets = d_env_offsets;
    p.n_own = n_own; p.n_opp = n_opp; p.E = n_envs; p.mode = mode; p.opt = g_opt();
    p.ept = 128 / (n_own > n_opp ? n_own : n_opp);
    p.n_tiles = (n_envs + p.ept - 1) / p.ept;
    // every checkpoint may own up to all tiles: one CTA per SM, at least one per checkpoint
    int grid = p.n_tiles * n_ckpt < sms ? p.n_tiles * n_ckpt : sms;
    if (grid < n_ckpt) grid = n_ckpt;
    mp::mp_policy_kernel<<<grid, mp::THREADS, mp::SMEM_BYTES, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "mp_forward_ensemble: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int mp_kernel_info(int n_own, int n_opp, int32_t *regs, int32_t *block, int32_t *smem, int32_t *blocks_per_sm,
                              int32_t *envs_per_tile) {
    if (n_own < 1 || n_own > MP_MAX_TEAM || n_opp < 1 || n_opp > MP_MAX_TEAM) return fa_internal_fail(-1, "mp_kernel_info: team sizes");
    if (int rc = prepare(nullptr)) return rc;
    cudaFuncAttributes at;
    cudaError_t e = cudaFuncGetAttributes(&at, mp::mp_policy_kernel);
    if (e != cudaSuccess) return fa_internal_fail(-2, "mp_kernel_info: %s", cudaGetErrorString(e));
    int nb = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, mp::mp_policy_kernel, mp::THREADS, mp::SMEM_BYTES);
    if (regs) *regs = at.numRegs;
    if (block) *block = mp::THREADS;
    if (smem) *smem = mp::SMEM_BYTES + (int)at.sharedSizeBytes;
    if (blocks_per_sm) *blocks_per_sm = nb;
    if (envs_per_tile) *envs_per_tile = 128 / (n_own > n_opp ? n_own : n_opp);
    return 0;
}

#ifdef MP_TRACE
// Diagnostic build (make -C csrc trace -> libfortattack_b200_trace.so): device buffer of 96 uint64 that receives clock64()
// of one row thread at every phase boundary of one tile of CTA 0 (NULL switches it off).  Not in the product library.
extern "C" int mp_set_trace(unsigned long long *d_trace) {
    g_trace = d_trace;
    g_trace_tile = 0;
    if (const char *ev = getenv("MP_TRACE_TILE")) g_trace_tile = atoi(ev);
    return 0;
}
#endif
