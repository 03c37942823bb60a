// rl_kernels.cu -- rollout-storage kernels (include/fortattack_rollout.h).
//
// rl_gae: one thread per (agent, env) column walks the T steps backwards; a warp's 32 lanes are 32
// consecutive envs of one agent, so every load/store is a coalesced 128-byte request.  HBM-bound streaming:
// 4 floats read (reward, V[t+1] is carried in a register, V[t], mask) + 1 written per (t, agent, env).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/fortattack_rollout.h"

int fa_internal_fail(int code, const char *fmt, ...);

namespace rl {

__global__ void __launch_bounds__(128) gae_kernel(const float *__restrict__ rew, float *__restrict__ val,
                                                  const float *__restrict__ next_value, const float *__restrict__ msk,
                                                  const uint8_t *__restrict__ ends, float *__restrict__ ret, int T, int A,
                                                  int E, float g, float gt) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (e >= E) return;
    const size_t plane = (size_t)A * E, col = (size_t)a * E + e;
    float v1 = next_value[col];                    // value_preds[T] = next_value (storage.py:61)
    val[(size_t)T * plane + col] = v1;
    float m1 = msk[(size_t)T * plane + col];
    bool end1 = ends[(size_t)T * E + e] != 0;
    float gae = 0.0f;
    for (int t = T - 1; t >= 0; --t) {
        const size_t k = (size_t)t * plane + col;
        const float r = rew[k], v0 = val[k], m0 = msk[k];
        const bool end0 = ends[(size_t)t * E + e] != 0;
        if (end1) gae = 0.0f;                                               // a segment ends right after step t
        // delta = r + gamma * V[t+1] * m[t+1] - V[t];  gae = delta + gamma * tau * m[t+1] * gae  (no FMA contraction)
        const float delta = __fsub_rn(__fadd_rn(r, __fmul_rn(__fmul_rn(g, v1), m1)), v0);
        gae = __fadd_rn(delta, __fmul_rn(__fmul_rn(gt, m1), gae));
        if (!end0) ret[k] = __fadd_rn(gae, v0);                             // index `end` itself is skipped
        else gae = 0.0f;
        v1 = v0; m1 = m0; end1 = end0;
    }
}

// ---- per-step rollout bookkeeping: masks of slot t+1, end points, episode reward sums --------------------------------
__global__ void __launch_bounds__(256) bookkeeping_kernel(const float *__restrict__ obs_t, const float *__restrict__ obs_t1,
                                                          const uint8_t *__restrict__ done, const float *__restrict__ rew,
                                                          float *__restrict__ masks_t1, uint8_t *__restrict__ ends_t1,
                                                          float *__restrict__ ep_rew, int A, int E) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x, a = blockIdx.y;
    if (e >= E) return;
    const size_t k = (size_t)a * E + e;
    const float alive_before = obs_t[k * 6];
    const bool fin = done[e] != 0;
    masks_t1[k] = fin ? obs_t1[k * 6] : alive_before;
    ep_rew[k] += rew[k] * alive_before;
    if (a == 0) ends_t1[e] = fin ? 1 : 0;
}

// ---- minibatch gather: one thread per output row (k, j); rows of one agent are consecutive threads ---------------
struct GatherParams {
    const int64_t *idx;
    int mb, T, A, E, a0, n, o0, m;
    const float *obs, *value_preds, *returns, *masks, *old_logp, *adv;
    const int64_t *actions;
    float *obs_own, *obs_opp, *o_value_preds, *o_returns, *o_masks, *o_old_logp, *o_adv, *alive, *alive_sum;
    int64_t *o_actions;
};

__global__ void __launch_bounds__(256) gather_kernel(const GatherParams p) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, k = blockIdx.y;     // sample j of team slot k
    float alive = 0.0f;
    if (j < p.mb) {
        const int64_t id = p.idx[j];
        const int t = (int)(id / p.E), e = (int)(id - (int64_t)t * p.E);
        const size_t plane = (size_t)p.A * p.E;
        if (k < p.n) {
            const size_t src = (size_t)t * plane + (size_t)(p.a0 + k) * p.E + e, dst = (size_t)k * p.mb + j;
            const float2 *o = reinterpret_cast<const float2 *>(p.obs + src * 6);
            float2 *d = reinterpret_cast<float2 *>(p.obs_own + dst * 6);
            const float2 o0 = o[0];
            d[0] = o0; d[1] = o[1]; d[2] = o[2];
            alive = o0.x;
            p.alive[dst] = alive;
            p.o_actions[dst] = p.actions[src];
            p.o_value_preds[dst] = p.value_preds[src];
            p.o_returns[dst] = p.returns[src];
            p.o_masks[dst] = p.masks[src];
            p.o_old_logp[dst] = p.old_logp[src];
            p.o_adv[dst] = p.adv[((size_t)k * p.T + t) * p.E + e];
        } else {
            const int ko = k - p.n;
            const size_t src = (size_t)t * plane + (size_t)(p.o0 + ko) * p.E + e, dst = (size_t)ko * p.mb + j;
            const float2 *o = reinterpret_cast<const float2 *>(p.obs + src * 6);
            float2 *d = reinterpret_cast<float2 *>(p.obs_opp + dst * 6);
            d[0] = o[0]; d[1] = o[1]; d[2] = o[2];
        }
    }
    // alive count of the block -> one atomic (the loss normaliser mask.sum())
    __shared__ float red[8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) alive += __shfl_xor_sync(0xffffffffu, alive, s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = alive;
    __syncthreads();
    if (threadIdx.x == 0 && k < p.n) {
        float t = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
        if (t != 0.0f) atomicAdd(p.alive_sum, t);
    }
}

// ---- masked clipped-PPO loss + gradient -------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ppo_loss_kernel(const float *__restrict__ values, const float *__restrict__ logp,
                                                       const float *__restrict__ entropy, const float *__restrict__ old_values,
                                                       const float *__restrict__ returns, const float *__restrict__ old_logp,
                                                       const float *__restrict__ adv, const float *__restrict__ mask,
                                                       const float *__restrict__ norm, int N, float clip, float vcoef, float ecoef,
                                                       float *out, float *gvalues, float *glogp, float *gentropy) {
    const float inv = 1.0f / norm[0];
    float sv = 0.0f, sa = 0.0f, se = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float m = mask[i], v = values[i], ov = old_values[i], ret = returns[i], a = adv[i];
        // entropy
        se += entropy[i] * m;
        gentropy[i] = -ecoef * m * inv;
        // surrogate
        const float ratio = m * expf(logp[i] - old_logp[i]);
        const float clamped = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
        const float s1 = ratio * a, s2 = clamped * a;
        sa += m * -fminf(s1, s2);
        const float d1 = ratio * a, d2 = (ratio >= 1.0f - clip && ratio <= 1.0f + clip) ? ratio * a : 0.0f;
        const float dmin = s1 < s2 ? d1 : (s1 > s2 ? d2 : 0.5f * (d1 + d2));
        glogp[i] = -m * dmin * inv;
        // clipped value loss
        const float dv = v - ov;
        const float vc = ov + fminf(fmaxf(dv, -clip), clip);
        const float e1 = v - ret, e2 = vc - ret;
        const float A1 = e1 * e1, A2 = e2 * e2;
        sv += 0.5f * fmaxf(A1, A2) * m;
        const float g1 = 2.0f * e1, g2 = (dv >= -clip && dv <= clip) ? 2.0f * e2 : 0.0f;
        const float dmax = A1 > A2 ? g1 : (A1 < A2 ? g2 : 0.5f * (g1 + g2));
        gvalues[i] = vcoef * 0.5f * dmax * m * inv;
    }
    __shared__ float red[3][8];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, s);
        sa += __shfl_xor_sync(0xffffffffu, sa, s);
        se += __shfl_xor_sync(0xffffffffu, se, s);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sa; red[2][threadIdx.x >> 5] = se; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tv = 0.0f, ta = 0.0f, te = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tv += red[0][w]; ta += red[1][w]; te += red[2][w]; }
        tv *= inv; ta *= inv; te *= inv;
        atomicAdd(out + 0, tv);
        atomicAdd(out + 1, ta);
        atomicAdd(out + 2, te);
        atomicAdd(out + 3, tv * vcoef + ta - te * ecoef);
    }
}

// ---- the same loss taken from the action head's LOGITS: Categorical log-prob / entropy (rlcore/distributions.py:9-17,
// mpnn.py:199-200) and their gradient folded in, per-block partial sums added in block order by the last block ----------
constexpr int LOSS_ACTIONS = 8;
__global__ void __launch_bounds__(256) ppo_loss_logits_kernel(const float *__restrict__ values, const float *__restrict__ logits,
                                                              const long long *__restrict__ actions, const float *__restrict__ old_values,
                                                              const float *__restrict__ returns, const float *__restrict__ old_logp,
                                                              const float *__restrict__ adv, const float *__restrict__ mask,
                                                              const float *__restrict__ norm, int N, float clip, float vcoef, float ecoef,
                                                              float *out, float *gvalues, float *glogits, float *logp_out,
                                                              float *entropy_out, float *scratch) {
    const float inv = norm != nullptr ? 1.0f / norm[0] : 1.0f;
    float sv = 0.0f, sa = 0.0f, se = 0.0f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < N; i += gridDim.x * blockDim.x) {
        const float4 l0 = *reinterpret_cast<const float4 *>(logits + (size_t)i * LOSS_ACTIONS);
        const float4 l1 = *reinterpret_cast<const float4 *>(logits + (size_t)i * LOSS_ACTIONS + 4);
        const float lg[LOSS_ACTIONS] = {l0.x, l0.y, l0.z, l0.w, l1.x, l1.y, l1.z, l1.w};
        float mx = lg[0];
#pragma unroll
        for (int k = 1; k < LOSS_ACTIONS; ++k) mx = fmaxf(mx, lg[k]);
        float pr[LOSS_ACTIONS], sum = 0.0f;
#pragma unroll
        for (int k = 0; k < LOSS_ACTIONS; ++k) { pr[k] = expf(lg[k] - mx); sum += pr[k]; }
        const float lse = mx + logf(sum), rs = 1.0f / sum;
        int act = (int)actions[i];
        act = act < 0 ? 0 : (act > LOSS_ACTIONS - 1 ? LOSS_ACTIONS - 1 : act);
        float lp = 0.0f, ent = 0.0f, ln[LOSS_ACTIONS];
#pragma unroll
        for (int k = 0; k < LOSS_ACTIONS; ++k) {
            pr[k] *= rs;
            ln[k] = lg[k] - lse;
            ent -= pr[k] * ln[k];
            lp = k == act ? ln[k] : lp;
        }
        if (logp_out != nullptr) logp_out[i] = lp;
        if (entropy_out != nullptr) entropy_out[i] = ent;
        if (out == nullptr) continue;                          // evaluation only (the behaviour log-probs of recompute_old)
        const float m = mask[i], v = values[i], ov = old_values[i], ret = returns[i], a = adv[i];
        se += ent * m;
        const float gen = -ecoef * m * inv;
        const float ratio = m * expf(lp - old_logp[i]);
        const float clamped = fminf(fmaxf(ratio, 1.0f - clip), 1.0f + clip);
        const float s1 = ratio * a, s2 = clamped * a;
        sa += m * -fminf(s1, s2);
        const float d1 = ratio * a, d2 = (ratio >= 1.0f - clip && ratio <= 1.0f + clip) ? ratio * a : 0.0f;
        const float dmin = s1 < s2 ? d1 : (s1 > s2 ? d2 : 0.5f * (d1 + d2));
        const float glp = -m * dmin * inv;
        // d logp_a / d l_j = [j == a] - p_j ;  d H / d l_j = -p_j (log p_j + H)
        float g[LOSS_ACTIONS];
#pragma unroll
        for (int k = 0; k < LOSS_ACTIONS; ++k) g[k] = glp * ((k == act ? 1.0f : 0.0f) - pr[k]) - gen * pr[k] * (ln[k] + ent);
        *reinterpret_cast<float4 *>(glogits + (size_t)i * LOSS_ACTIONS) = make_float4(g[0], g[1], g[2], g[3]);
        *reinterpret_cast<float4 *>(glogits + (size_t)i * LOSS_ACTIONS + 4) = make_float4(g[4], g[5], g[6], g[7]);
        const float dv = v - ov;
        const float vc = ov + fminf(fmaxf(dv, -clip), clip);
        const float e1 = v - ret, e2 = vc - ret;
        const float A1 = e1 * e1, A2 = e2 * e2;
        sv += 0.5f * fmaxf(A1, A2) * m;
        const float g1 = 2.0f * e1, g2 = (dv >= -clip && dv <= clip) ? 2.0f * e2 : 0.0f;
        const float dmax = A1 > A2 ? g1 : (A1 < A2 ? g2 : 0.5f * (g1 + g2));
        gvalues[i] = vcoef * 0.5f * dmax * m * inv;
    }
    if (out == nullptr) return;
    __shared__ float red[3][8];
    __shared__ bool last;
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, s);
        sa += __shfl_xor_sync(0xffffffffu, sa, s);
        se += __shfl_xor_sync(0xffffffffu, se, s);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sa; red[2][threadIdx.x >> 5] = se; }
    __syncthreads();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(scratch);
    if (threadIdx.x == 0) {
        float tv = 0.0f, ta = 0.0f, te = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tv += red[0][w]; ta += red[1][w]; te += red[2][w]; }
        float *mine = scratch + 4 + 3 * blockIdx.x;
        mine[0] = tv; mine[1] = ta; mine[2] = te;
        __threadfence();
        last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!last) return;
    // the last block to arrive adds the per-block partials in a FIXED order (thread t: blocks t, t + 256, ...; then the same
    // shuffle / warp-order tree as above), so the statistics do not depend on the order the blocks finished in
    __threadfence();
    const volatile float *part = scratch + 4;
    sv = sa = se = 0.0f;
    for (unsigned int b = threadIdx.x; b < gridDim.x; b += blockDim.x) { sv += part[3 * b]; sa += part[3 * b + 1]; se += part[3 * b + 2]; }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) {
        sv += __shfl_xor_sync(0xffffffffu, sv, s);
        sa += __shfl_xor_sync(0xffffffffu, sa, s);
        se += __shfl_xor_sync(0xffffffffu, se, s);
    }
    __syncthreads();                                           // (red is reused)
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sa; red[2][threadIdx.x >> 5] = se; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float tv = 0.0f, ta = 0.0f, te = 0.0f;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) { tv += red[0][w]; ta += red[1][w]; te += red[2][w]; }
        tv *= inv; ta *= inv; te *= inv;
        out[0] = tv; out[1] = ta; out[2] = te; out[3] = tv * vcoef + ta - te * ecoef;
        *ticket = 0u;
    }
}

// ---- tiny attention: one warp per batch element, the 32 lanes split the feature dimension (VEC floats each) --------
constexpr int ATT_MAX = 5;
struct Opnd { float *p; long long bs, rs; };

template <int VEC> struct VecLoad;
template <> struct VecLoad<1> { static __device__ __forceinline__ void ld(const float *p, float (&v)[1]) { v[0] = *p; }
                                static __device__ __forceinline__ void st(float *p, const float (&v)[1]) { *p = v[0]; } };
template <> struct VecLoad<2> { static __device__ __forceinline__ void ld(const float *p, float (&v)[2]) { const float2 t = *reinterpret_cast<const float2 *>(p); v[0] = t.x; v[1] = t.y; }
                                static __device__ __forceinline__ void st(float *p, const float (&v)[2]) { *reinterpret_cast<float2 *>(p) = make_float2(v[0], v[1]); } };
template <> struct VecLoad<3> { static __device__ __forceinline__ void ld(const float *p, float (&v)[3]) { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; }
                                static __device__ __forceinline__ void st(float *p, const float (&v)[3]) { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; } };
template <> struct VecLoad<4> { static __device__ __forceinline__ void ld(const float *p, float (&v)[4]) { const float4 t = *reinterpret_cast<const float4 *>(p); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
                                static __device__ __forceinline__ void st(float *p, const float (&v)[4]) { *reinterpret_cast<float4 *>(p) = make_float4(v[0], v[1], v[2], v[3]); } };

__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) x += __shfl_xor_sync(0xffffffffu, x, s);
    return x;
}

// MIX = keys and values are the same rows X (rl_attn_mix_*): X is loaded once, the backward accumulates ONE gradient for it.
template <int VEC, bool MIX>
__global__ void __launch_bounds__(128) attn_fwd_kernel(Opnd A, Opnd B, Opnd V, Opnd O, float *attn, int batch, int n, int m,
                                                       float norm, int mask_diag, Opnd C) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const int c = lane * VEC;
    float bb[ATT_MAX][VEC], vvs[MIX ? 1 : ATT_MAX][VEC];
    float (&vv)[ATT_MAX][VEC] = *reinterpret_cast<float (*)[ATT_MAX][VEC]>(MIX ? &bb[0][0] : &vvs[0][0]);
#pragma unroll
    for (int j = 0; j < ATT_MAX; ++j)
        if (j < m) {
            VecLoad<VEC>::ld(B.p + b * B.bs + j * B.rs + c, bb[j]);
            if (!MIX) VecLoad<VEC>::ld(V.p + b * V.bs + j * V.rs + c, vv[j]);
            if (MIX && C.p) VecLoad<VEC>::st(C.p + b * C.bs + j * C.rs + c, bb[j]);     // the X rows, re-emitted
        }
#pragma unroll
    for (int i = 0; i < ATT_MAX; ++i) {
        if (i < n) {
            float a[VEC];
            VecLoad<VEC>::ld(A.p + b * A.bs + i * A.rs + c, a);
            float s[ATT_MAX], mx = -INFINITY;
#pragma unroll
            for (int j = 0; j < ATT_MAX; ++j) {
                s[j] = -INFINITY;
                if (j < m) {
                    float d = 0.0f;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) d = fmaf(a[q], bb[j][q], d);
                    d = warp_sum(d) * norm;
                    s[j] = (mask_diag && i == j) ? -INFINITY : d;
                    mx = fmaxf(mx, s[j]);
                }
            }
            float sum = 0.0f;
#pragma unroll
            for (int j = 0; j < ATT_MAX; ++j) {
                s[j] = (j < m && s[j] != -INFINITY) ? expf(s[j] - mx) : 0.0f;
                sum += s[j];
            }
            const float inv = sum > 0.0f ? 1.0f / sum : 0.0f;
            float o[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) o[q] = 0.0f;
#pragma unroll
            for (int j = 0; j < ATT_MAX; ++j)
                if (j < m) {
                    s[j] *= inv;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) o[q] = fmaf(s[j], vv[j][q], o[q]);
                    if (lane == j) attn[(b * n + i) * m + j] = s[j];
                }
            VecLoad<VEC>::st(O.p + b * O.bs + i * O.rs + c, o);
        }
    }
}

// MIX: dB receives dB + dV (+ Cin, a gradient that reaches X by another path; may be null); dV is not written.
template <int VEC, bool MIX>
__global__ void __launch_bounds__(128) attn_bwd_kernel(Opnd G, Opnd A, Opnd B, Opnd V, const float *attn, Opnd dA, Opnd dB,
                                                       Opnd dV, int batch, int n, int m, float norm, Opnd Cin) {
    const int lane = threadIdx.x & 31;
    const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (b >= batch) return;
    const int c = lane * VEC;
    float bb[ATT_MAX][VEC], db[ATT_MAX][VEC], vvs[MIX ? 1 : ATT_MAX][VEC], dvs[MIX ? 1 : ATT_MAX][VEC];
    float (&vv)[ATT_MAX][VEC] = *reinterpret_cast<float (*)[ATT_MAX][VEC]>(MIX ? &bb[0][0] : &vvs[0][0]);
    float (&dv)[ATT_MAX][VEC] = *reinterpret_cast<float (*)[ATT_MAX][VEC]>(MIX ? &db[0][0] : &dvs[0][0]);
#pragma unroll
    for (int j = 0; j < ATT_MAX; ++j)
        if (j < m) {
            VecLoad<VEC>::ld(B.p + b * B.bs + j * B.rs + c, bb[j]);
            if (!MIX) VecLoad<VEC>::ld(V.p + b * V.bs + j * V.rs + c, vv[j]);
#pragma unroll
            for (int q = 0; q < VEC; ++q) { db[j][q] = 0.0f; dv[j][q] = 0.0f; }
            if (MIX && Cin.p) VecLoad<VEC>::ld(Cin.p + b * Cin.bs + j * Cin.rs + c, db[j]);
        }
#pragma unroll
    for (int i = 0; i < ATT_MAX; ++i) {
        if (i < n) {
            float a[VEC], g[VEC], p[ATT_MAX], dp[ATT_MAX];
            VecLoad<VEC>::ld(A.p + b * A.bs + i * A.rs + c, a);
            VecLoad<VEC>::ld(G.p + b * G.bs + i * G.rs + c, g);
            float dot = 0.0f;                       // sum_j p_ij dp_ij
#pragma unroll
            for (int j = 0; j < ATT_MAX; ++j) {
                p[j] = 0.0f; dp[j] = 0.0f;
                if (j < m) {
                    p[j] = attn[(b * n + i) * m + j];
                    float d = 0.0f;
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        d = fmaf(g[q], vv[j][q], d);
                        dv[j][q] = fmaf(p[j], g[q], dv[j][q]);
                    }
                    dp[j] = warp_sum(d);
                    dot = fmaf(p[j], dp[j], dot);
                }
            }
            float da[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) da[q] = 0.0f;
#pragma unroll
            for (int j = 0; j < ATT_MAX; ++j)
                if (j < m) {
                    const float ds = p[j] * (dp[j] - dot) * norm;     // softmax backward; masked entries have p = 0
#pragma unroll
                    for (int q = 0; q < VEC; ++q) {
                        da[q] = fmaf(ds, bb[j][q], da[q]);
                        db[j][q] = fmaf(ds, a[q], db[j][q]);
                    }
                }
            VecLoad<VEC>::st(dA.p + b * dA.bs + i * dA.rs + c, da);
        }
    }
#pragma unroll
    for (int j = 0; j < ATT_MAX; ++j)
        if (j < m) {
            VecLoad<VEC>::st(dB.p + b * dB.bs + j * dB.rs + c, db[j]);
            if (!MIX) VecLoad<VEC>::st(dV.p + b * dV.bs + j * dV.rs + c, dv[j]);
        }
}

// ---- ReLU backward fused with the bias gradient: dpre = dout * [out > 0], partial[block][c] = sum over the block's rows
// of dpre[., c].  Row-major [rows, cols], cols = 4 * c4 with c4 in {8, 16, 32, 64}: c4 lanes cover one row with 16-byte
// accesses, 256 / c4 rows per block iteration, a fixed block -> rows assignment and a fixed summation order (the
// caller adds the [blocks, cols] partials), so the result is reproducible run to run.
__global__ void __launch_bounds__(256) relu_bwd_colsum_kernel(const float4 *__restrict__ dout, const float4 *__restrict__ out,
                                                              float4 *__restrict__ dpre, float4 *__restrict__ partial,
                                                              long long rows, int c4, long long ldd4, long long ldo4, long long ldp4) {
    __shared__ float4 sm[256];
    const int lane = threadIdx.x % c4, rsub = threadIdx.x / c4, rpi = 256 / c4;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (long long r = (long long)blockIdx.x * rpi + rsub; r < rows; r += (long long)gridDim.x * rpi) {
        float4 g = dout[r * ldd4 + lane];
        const float4 o = out[r * ldo4 + lane];
        g.x = o.x > 0.f ? g.x : 0.f; g.y = o.y > 0.f ? g.y : 0.f; g.z = o.z > 0.f ? g.z : 0.f; g.w = o.w > 0.f ? g.w : 0.f;
        dpre[r * ldp4 + lane] = g;
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    if (rsub == 0) {
        for (int q = 1; q < rpi; ++q) {
            const float4 t = sm[q * c4 + lane];
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        partial[(long long)blockIdx.x * c4 + lane] = acc;
    }
}

// ---- column sums of a tall matrix (bias gradients: sum over the rows of dpre, or of relu_bwd_colsum's per-block partials) ----
// cols is a power of two <= 256: thread t owns column t % cols and every (256 / cols)-th row of the block's row range; the
// row lanes of a column are added in lane order, the blocks' partials in block order by the last block to arrive.
__global__ void __launch_bounds__(256) colsum_kernel(const float *__restrict__ x, long long rows, int cols, long long ld,
                                                     float *__restrict__ out, float *scratch) {
    __shared__ float sm[256];
    __shared__ bool last;
    const int c = threadIdx.x % cols, lane = threadIdx.x / cols, lanes = 256 / cols;
    const long long per = (rows + gridDim.x - 1) / gridDim.x, r0 = (long long)blockIdx.x * per;
    const long long r1 = r0 + per < rows ? r0 + per : rows;
    float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
    long long r = r0 + lane;
    for (; r + 3ll * lanes < r1; r += 4ll * lanes) {
        a0 += x[r * ld + c]; a1 += x[(r + lanes) * ld + c]; a2 += x[(r + 2ll * lanes) * ld + c]; a3 += x[(r + 3ll * lanes) * ld + c];
    }
    for (; r < r1; r += lanes) a0 += x[r * ld + c];
    sm[threadIdx.x] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    unsigned int *ticket = reinterpret_cast<unsigned int *>(scratch);
    float *part = scratch + 4;
    if (lane == 0) {
        float t = sm[c];
        for (int q = 1; q < lanes; ++q) t += sm[q * cols + c];
        part[(size_t)blockIdx.x * cols + c] = t;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    __syncthreads();
    if (!last) return;
    __threadfence();
    const volatile float *vp = part;
    float t = 0.0f;
    for (unsigned int b = lane; b < gridDim.x; b += lanes) t += vp[(size_t)b * cols + c];
    __syncthreads();
    sm[threadIdx.x] = t;
    __syncthreads();
    if (lane == 0) {
        t = sm[c];
        for (int q = 1; q < lanes; ++q) t += sm[q * cols + c];
        out[c] = t;
    }
    if (threadIdx.x == 0) *ticket = 0u;
}

// ---- a handful of small fp32 products in one launch (the [d, d] folds of the attention projections and their backward) ----
// C (+)= op(A) op(B), every dimension <= 256: 32 x 32 output tiles, one per block, K in steps of 32 through shared memory; plain
// fp32 FMAs in a fixed order (bit-reproducible).  The folds are ~2 MFLOP each: what they cost before was their LAUNCHES (one
// persistent tcgen05 launch + a weight pack + a partial-sum reduction per product).
struct SmallMM { RlSmallMatmul it[RL_SMALL_MATMUL_MAX]; int tile0[RL_SMALL_MATMUL_MAX + 1]; int n; };

__global__ void __launch_bounds__(256) small_matmul_kernel(const SmallMM p) {
    __shared__ float As[32][33], Bs[32][33];
    int k = 0;
    while (k + 1 < p.n && (int)blockIdx.x >= p.tile0[k + 1]) ++k;
    const RlSmallMatmul m = p.it[k];
    const int t = (int)blockIdx.x - p.tile0[k], tn = (m.N + 31) / 32, i0 = (t / tn) * 32, j0 = (t % tn) * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;           // 8 rows of threads: thread (ty, tx) -> outputs (i0 + ty + 8 q, j0 + tx)
    float acc[4] = {0.f, 0.f, 0.f, 0.f};
    for (int k0 = 0; k0 < m.K; k0 += 32) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = ty + 8 * q;
            // As[r][c] = op(A)[i0 + r][k0 + c],  Bs[r][c] = op(B)[k0 + r][j0 + c]; tx always runs along the dimension that is
            // contiguous in memory (full-line loads), transposed operands are turned through the padded tiles
            if (m.trans_a) {
                const int ak = k0 + r, ai = i0 + tx;
                As[tx][r] = (ai < m.M && ak < m.K) ? m.A[(size_t)ak * m.lda + ai] : 0.0f;
            } else {
                const int ai = i0 + r, ak = k0 + tx;
                As[r][tx] = (ai < m.M && ak < m.K) ? m.A[(size_t)ai * m.lda + ak] : 0.0f;
            }
            if (m.trans_b) {
                const int bj = j0 + r, bk = k0 + tx;
                Bs[tx][r] = (bk < m.K && bj < m.N) ? m.B[(size_t)bj * m.ldb + bk] : 0.0f;
            } else {
                const int bk = k0 + r, bj = j0 + tx;
                Bs[r][tx] = (bk < m.K && bj < m.N) ? m.B[(size_t)bk * m.ldb + bj] : 0.0f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
            const float b = Bs[c][tx];
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaf(As[ty + 8 * q][c], b, acc[q]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const int i = i0 + ty + 8 * q, j = j0 + tx;
        if (i < m.M && j < m.N) {
            float *c = m.C + (size_t)i * m.ldc + j;
            *c = m.accumulate ? *c + acc[q] : acc[q];
        }
    }
}

}  // namespace rl

static int attn_check(const char *who, int batch, int n, int m, int k, const RlAttnOperand *const *ops, int nops) {
    if (batch < 1 || n < 1 || m < 1 || n > rl::ATT_MAX || m > rl::ATT_MAX || k < 32 || k > 128 || k % 32 != 0)
        return fa_internal_fail(-1, "%s: need 1 <= n, m <= %d and k in {32, 64, 96, 128}", who, rl::ATT_MAX);
    const int vec = k / 32;
    for (int i = 0; i < nops; ++i) {
        if (!ops[i] || !ops[i]->ptr) return fa_internal_fail(-1, "%s: NULL operand", who);
        if (vec == 2 || vec == 4) {     // vector loads: every row must start on a VEC * 4 byte boundary
            const uintptr_t al = (uintptr_t)vec * 4;
            if (((uintptr_t)ops[i]->ptr % al) || (ops[i]->batch_stride % vec) || (ops[i]->row_stride % vec))
                return fa_internal_fail(-4, "%s: operand %d is not aligned for %d-float vector access", who, i, vec);
        }
    }
    return 0;
}
static const rl::Opnd NOOP{nullptr, 0, 0};
static rl::Opnd opnd(const RlAttnOperand *o) { return rl::Opnd{o->ptr, (long long)o->batch_stride, (long long)o->row_stride}; }

extern "C" int rl_attn_forward(const RlAttnOperand *A, const RlAttnOperand *B, const RlAttnOperand *V, const RlAttnOperand *out,
                               float *d_attn, int batch, int n, int m, int k, float norm, int mask_diag, void *stream) {
    const RlAttnOperand *ops[4] = {A, B, V, out};
    if (int rc = attn_check("rl_attn_forward", batch, n, m, k, ops, 4)) return rc;
    if (!d_attn) return fa_internal_fail(-1, "rl_attn_forward: NULL attention output");
    const int grid = (batch + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    switch (k / 32) {
        case 1: rl::attn_fwd_kernel<1, false><<<grid, 128, 0, st>>>(opnd(A), opnd(B), opnd(V), opnd(out), d_attn, batch, n, m, norm, mask_diag, NOOP); break;
        case 2: rl::attn_fwd_kernel<2, false><<<grid, 128, 0, st>>>(opnd(A), opnd(B), opnd(V), opnd(out), d_attn, batch, n, m, norm, mask_diag, NOOP); break;
        case 3: rl::attn_fwd_kernel<3, false><<<grid, 128, 0, st>>>(opnd(A), opnd(B), opnd(V), opnd(out), d_attn, batch, n, m, norm, mask_diag, NOOP); break;
        default: rl::attn_fwd_kernel<4, false><<<grid, 128, 0, st>>>(opnd(A), opnd(B), opnd(V), opnd(out), d_attn, batch, n, m, norm, mask_diag, NOOP); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_attn_forward: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_attn_backward(const RlAttnOperand *dout, const RlAttnOperand *A, const RlAttnOperand *B, const RlAttnOperand *V,
                                const float *d_attn, const RlAttnOperand *dA, const RlAttnOperand *dB, const RlAttnOperand *dV,
                                int batch, int n, int m, int k, float norm, void *stream) {
    const RlAttnOperand *ops[7] = {dout, A, B, V, dA, dB, dV};
    if (int rc = attn_check("rl_attn_backward", batch, n, m, k, ops, 7)) return rc;
    if (!d_attn) return fa_internal_fail(-1, "rl_attn_backward: NULL attention input");
    const int grid = (batch + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    switch (k / 32) {
        case 1: rl::attn_bwd_kernel<1, false><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(B), opnd(V), d_attn, opnd(dA), opnd(dB), opnd(dV), batch, n, m, norm, NOOP); break;
        case 2: rl::attn_bwd_kernel<2, false><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(B), opnd(V), d_attn, opnd(dA), opnd(dB), opnd(dV), batch, n, m, norm, NOOP); break;
        case 3: rl::attn_bwd_kernel<3, false><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(B), opnd(V), d_attn, opnd(dA), opnd(dB), opnd(dV), batch, n, m, norm, NOOP); break;
        default: rl::attn_bwd_kernel<4, false><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(B), opnd(V), d_attn, opnd(dA), opnd(dB), opnd(dV), batch, n, m, norm, NOOP); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_attn_backward: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_attn_mix_forward(const RlAttnOperand *A, const RlAttnOperand *X, const RlAttnOperand *out,
                                   const RlAttnOperand *x_copy, float *d_attn, int batch, int n, int m, int k, float norm,
                                   int mask_diag, void *stream) {
    const RlAttnOperand *ops[4] = {A, X, out, x_copy ? x_copy : X};
    if (int rc = attn_check("rl_attn_mix_forward", batch, n, m, k, ops, 4)) return rc;
    if (!d_attn) return fa_internal_fail(-1, "rl_attn_mix_forward: NULL attention output");
    const int grid = (batch + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    const rl::Opnd C = x_copy ? opnd(x_copy) : rl::Opnd{nullptr, 0, 0};
    switch (k / 32) {
        case 1: rl::attn_fwd_kernel<1, true><<<grid, 128, 0, st>>>(opnd(A), opnd(X), opnd(X), opnd(out), d_attn, batch, n, m, norm, mask_diag, C); break;
        case 2: rl::attn_fwd_kernel<2, true><<<grid, 128, 0, st>>>(opnd(A), opnd(X), opnd(X), opnd(out), d_attn, batch, n, m, norm, mask_diag, C); break;
        case 3: rl::attn_fwd_kernel<3, true><<<grid, 128, 0, st>>>(opnd(A), opnd(X), opnd(X), opnd(out), d_attn, batch, n, m, norm, mask_diag, C); break;
        default: rl::attn_fwd_kernel<4, true><<<grid, 128, 0, st>>>(opnd(A), opnd(X), opnd(X), opnd(out), d_attn, batch, n, m, norm, mask_diag, C); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_attn_mix_forward: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_attn_mix_backward(const RlAttnOperand *dout, const RlAttnOperand *A, const RlAttnOperand *X, const float *d_attn,
                                    const RlAttnOperand *dA, const RlAttnOperand *dX, const RlAttnOperand *dx_add, int batch,
                                    int n, int m, int k, float norm, void *stream) {
    const RlAttnOperand *ops[6] = {dout, A, X, dA, dX, dx_add ? dx_add : dX};
    if (int rc = attn_check("rl_attn_mix_backward", batch, n, m, k, ops, 6)) return rc;
    if (!d_attn) return fa_internal_fail(-1, "rl_attn_mix_backward: NULL attention input");
    const int grid = (batch + 3) / 4;
    cudaStream_t st = (cudaStream_t)stream;
    const rl::Opnd C = dx_add ? opnd(dx_add) : rl::Opnd{nullptr, 0, 0};
    switch (k / 32) {
        case 1: rl::attn_bwd_kernel<1, true><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(X), opnd(X), d_attn, opnd(dA), opnd(dX), opnd(dX), batch, n, m, norm, C); break;
        case 2: rl::attn_bwd_kernel<2, true><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(X), opnd(X), d_attn, opnd(dA), opnd(dX), opnd(dX), batch, n, m, norm, C); break;
        case 3: rl::attn_bwd_kernel<3, true><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(X), opnd(X), d_attn, opnd(dA), opnd(dX), opnd(dX), batch, n, m, norm, C); break;
        default: rl::attn_bwd_kernel<4, true><<<grid, 128, 0, st>>>(opnd(dout), opnd(A), opnd(X), opnd(X), d_attn, opnd(dA), opnd(dX), opnd(dX), batch, n, m, norm, C); break;
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_attn_mix_backward: launch: %s", cudaGetErrorString(e));
    return 0;
}

// 8 resident blocks per SM of the current device (the count is a property of the chip, queried once per device; it only
// sizes grids -- without a device, e.g. in the CPU-side symbol tests, the B200's 148 is assumed)
static int grid_cap() {
    static int cap[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148 * 8;
    if (cap[dev] == 0) {
        int sms = 0;
        cap[dev] = (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && sms > 0 ? sms : 148) * 8;
    }
    return cap[dev];
}

extern "C" int rl_relu_bwd_colsum_blocks(long long rows, int cols) {
    if (rows < 1 || (cols != 32 && cols != 64 && cols != 128 && cols != 256)) return -1;
    const long long rpi = 256 / (cols / 4), need = (rows + rpi - 1) / rpi;
    const int cap = grid_cap();
    return (int)(need < cap ? need : cap);
}

extern "C" int rl_relu_bwd_colsum(const float *d_dout, const float *d_out, float *d_dpre, float *d_partial, long long rows,
                                  int cols, void *stream) {
    return rl_relu_bwd_colsum_ld(d_dout, cols, d_out, cols, d_dpre, cols, d_partial, rows, cols, stream);
}

extern "C" int rl_relu_bwd_colsum_ld(const float *d_dout, int ldd, const float *d_out, int ldo, float *d_dpre, int ldp,
                                     float *d_partial, long long rows, int cols, void *stream) {
    if (!d_dout || !d_out || !d_dpre || !d_partial) return fa_internal_fail(-1, "rl_relu_bwd_colsum: NULL pointer");
    const int blocks = rl_relu_bwd_colsum_blocks(rows, cols);
    if (blocks < 1) return fa_internal_fail(-1, "rl_relu_bwd_colsum: need rows >= 1 and cols in {32, 64, 128, 256} (got %lld x %d)", rows, cols);
    if (((uintptr_t)d_dout | (uintptr_t)d_out | (uintptr_t)d_dpre | (uintptr_t)d_partial) % 16)
        return fa_internal_fail(-4, "rl_relu_bwd_colsum: pointers must be 16-byte aligned");
    if (ldd < cols || ldo < cols || ldp < cols || (ldd | ldo | ldp) % 4)
        return fa_internal_fail(-1, "rl_relu_bwd_colsum: row strides must be >= cols and multiples of 4 floats");
    rl::relu_bwd_colsum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(d_dout), reinterpret_cast<const float4 *>(d_out), reinterpret_cast<float4 *>(d_dpre),
        reinterpret_cast<float4 *>(d_partial), rows, cols / 4, ldd / 4, ldo / 4, ldp / 4);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_relu_bwd_colsum: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_small_matmul(const RlSmallMatmul *items, int n_items, void *stream) {
    if (!items || n_items < 1 || n_items > RL_SMALL_MATMUL_MAX)
        return fa_internal_fail(-1, "rl_small_matmul: 1 <= n_items <= %d", RL_SMALL_MATMUL_MAX);
    rl::SmallMM p;
    p.n = n_items;
    int tiles = 0;
    for (int k = 0; k < n_items; ++k) {
        const RlSmallMatmul &m = items[k];
        if (!m.A || !m.B || !m.C || m.M < 1 || m.N < 1 || m.K < 1 || m.M > 256 || m.N > 256 || m.K > 256 ||
            m.lda < (m.trans_a ? m.M : m.K) || m.ldb < (m.trans_b ? m.K : m.N) || m.ldc < m.N)
            return fa_internal_fail(-1, "rl_small_matmul: item %d: NULL pointer, a dimension outside 1..256 or a row stride below the row length", k);
        p.it[k] = m;
        p.tile0[k] = tiles;
        tiles += ((m.M + 31) / 32) * ((m.N + 31) / 32);
    }
    p.tile0[n_items] = tiles;
    rl::small_matmul_kernel<<<tiles, 256, 0, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_small_matmul: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_colsum_blocks(long long rows, int cols) {
    if (rows < 1 || cols < 1 || cols > 256 || (cols & (cols - 1))) return -1;
    const long long lanes = 256 / cols, need = (rows + lanes * 16 - 1) / (lanes * 16);
    const int cap = 4 * grid_cap() < 1024 ? 4 * grid_cap() : 1024;
    return (int)(need < 1 ? 1 : (need < cap ? need : cap));
}

extern "C" int rl_colsum(const float *d_x, long long rows, int cols, int ld, float *d_out, float *d_scratch, void *stream) {
    if (!d_x || !d_out || !d_scratch) return fa_internal_fail(-1, "rl_colsum: NULL pointer");
    const int blocks = rl_colsum_blocks(rows, cols);
    if (blocks < 1 || ld < cols)
        return fa_internal_fail(-1, "rl_colsum: need rows >= 1, cols a power of two <= 256, ld >= cols (got %lld x %d, ld %d)", rows, cols, ld);
    rl::colsum_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_x, rows, cols, ld, d_out, d_scratch);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_colsum: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_rollout_bookkeeping(const float *d_obs_t, const float *d_obs_t1, const uint8_t *d_done, const float *d_reward,
                                      float *d_masks_t1, uint8_t *d_ends_t1, float *d_episode_rewards, int A, int E,
                                      void *stream) {
    if (!d_obs_t || !d_obs_t1 || !d_done || !d_reward || !d_masks_t1 || !d_ends_t1 || !d_episode_rewards)
        return fa_internal_fail(-1, "rl_rollout_bookkeeping: NULL pointer");
    if (A < 1 || E < 1 || A > 65535) return fa_internal_fail(-1, "rl_rollout_bookkeeping: bad sizes A=%d E=%d", A, E);
    const dim3 grid((unsigned)((E + 255) / 256), (unsigned)A);
    rl::bookkeeping_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(d_obs_t, d_obs_t1, d_done, d_reward, d_masks_t1, d_ends_t1,
                                                                   d_episode_rewards, A, E);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_rollout_bookkeeping: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_gather_minibatch(const int64_t *d_idx, int mb, int T, int A, int E, int a0, int n, int o0, int m,
                                   const float *d_obs, const int64_t *d_actions, const float *d_value_preds,
                                   const float *d_returns, const float *d_masks, const float *d_old_logp, const float *d_adv,
                                   float *obs_own, float *obs_opp, int64_t *actions, float *value_preds, float *returns,
                                   float *masks, float *old_logp, float *adv, float *alive, float *d_alive_sum, void *stream) {
    if (!d_idx || !d_obs || !d_actions || !d_value_preds || !d_returns || !d_masks || !d_old_logp || !d_adv || !obs_own ||
        !obs_opp || !actions || !value_preds || !returns || !masks || !old_logp || !adv || !alive || !d_alive_sum)
        return fa_internal_fail(-1, "rl_gather_minibatch: NULL pointer");
    if (mb < 1 || T < 1 || E < 1 || n < 1 || m < 1 || a0 < 0 || o0 < 0 || a0 + n > A || o0 + m > A)
        return fa_internal_fail(-1, "rl_gather_minibatch: bad sizes");
    rl::GatherParams p;
    p.idx = d_idx; p.mb = mb; p.T = T; p.A = A; p.E = E; p.a0 = a0; p.n = n; p.o0 = o0; p.m = m;
    p.obs = d_obs; p.actions = d_actions; p.value_preds = d_value_preds; p.returns = d_returns; p.masks = d_masks;
    p.old_logp = d_old_logp; p.adv = d_adv; p.obs_own = obs_own; p.obs_opp = obs_opp; p.o_actions = actions;
    p.o_value_preds = value_preds; p.o_returns = returns; p.o_masks = masks; p.o_old_logp = old_logp; p.o_adv = adv;
    p.alive = alive; p.alive_sum = d_alive_sum;
    const dim3 grid((unsigned)((mb + 255) / 256), (unsigned)(n + m));
    rl::gather_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_gather_minibatch: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_ppo_loss(const float *d_values, const float *d_logp, const float *d_entropy, const float *d_old_values,
                           const float *d_returns, const float *d_old_logp, const float *d_adv, const float *d_mask,
                           const float *d_norm, int N, float clip, float vcoef, float ecoef, float *d_out, float *d_gvalues,
                           float *d_glogp, float *d_gentropy, void *stream) {
    if (!d_values || !d_logp || !d_entropy || !d_old_values || !d_returns || !d_old_logp || !d_adv || !d_mask || !d_norm ||
        !d_out || !d_gvalues || !d_glogp || !d_gentropy)
        return fa_internal_fail(-1, "rl_ppo_loss: NULL pointer");
    if (N < 1) return fa_internal_fail(-1, "rl_ppo_loss: N must be >= 1");
    int blocks = (N + 255) / 256;
    if (blocks > grid_cap()) blocks = grid_cap();
    rl::ppo_loss_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_values, d_logp, d_entropy, d_old_values, d_returns,
                                                                  d_old_logp, d_adv, d_mask, d_norm, N, clip, vcoef, ecoef,
                                                                  d_out, d_gvalues, d_glogp, d_gentropy);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_ppo_loss: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" size_t rl_ppo_loss_logits_scratch_floats(void) { return (size_t)4 + 3 * (size_t)grid_cap(); }

extern "C" int rl_ppo_loss_logits(const float *d_values, const float *d_logits, const int64_t *d_actions, const float *d_old_values,
                                  const float *d_returns, const float *d_old_logp, const float *d_adv, const float *d_mask,
                                  const float *d_norm, int N, int n_actions, float clip, float vcoef, float ecoef, float *d_out,
                                  float *d_gvalues, float *d_glogits, float *d_logp, float *d_entropy, float *d_scratch,
                                  void *stream) {
    if (!d_logits || !d_actions) return fa_internal_fail(-1, "rl_ppo_loss_logits: NULL logits / actions");
    if (d_out != nullptr && (!d_values || !d_old_values || !d_returns || !d_old_logp || !d_adv || !d_mask || !d_gvalues ||
                             !d_glogits || !d_scratch))
        return fa_internal_fail(-1, "rl_ppo_loss_logits: NULL pointer (the loss form needs every input, both gradients and the scratch)");
    if (d_out == nullptr && !d_logp && !d_entropy) return fa_internal_fail(-1, "rl_ppo_loss_logits: nothing to compute");
    if (N < 1 || n_actions != rl::LOSS_ACTIONS)
        return fa_internal_fail(-1, "rl_ppo_loss_logits: N >= 1 and n_actions == %d expected (got %d, %d)", rl::LOSS_ACTIONS, N, n_actions);
    if (((uintptr_t)d_logits & 15) || ((uintptr_t)d_glogits & 15))
        return fa_internal_fail(-4, "rl_ppo_loss_logits: logits / gradient rows must be 16-byte aligned");
    int blocks = (N + 255) / 256;
    if (blocks > grid_cap()) blocks = grid_cap();
    rl::ppo_loss_logits_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        d_values, d_logits, (const long long *)d_actions, d_old_values, d_returns, d_old_logp, d_adv, d_mask, d_norm, N, clip, vcoef,
        ecoef, d_out, d_gvalues, d_glogits, d_logp, d_entropy, d_scratch);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_ppo_loss_logits: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int rl_gae(const float *d_rewards, float *d_value_preds, const float *d_next_value, const float *d_masks,
                      const uint8_t *d_ends, float *d_returns, int T, int A, int E, double gamma, double tau, void *stream) {
    if (!d_rewards || !d_value_preds || !d_next_value || !d_masks || !d_ends || !d_returns)
        return fa_internal_fail(-1, "rl_gae: NULL pointer");
    if (T < 1 || A < 1 || E < 1 || A > 65535) return fa_internal_fail(-1, "rl_gae: bad sizes T=%d A=%d E=%d", T, A, E);
    const dim3 grid((unsigned)((E + 127) / 128), (unsigned)A);
    rl::gae_kernel<<<grid, 128, 0, (cudaStream_t)stream>>>(d_rewards, d_value_preds, d_next_value, d_masks, d_ends, d_returns,
                                                           T, A, E, (float)gamma, (float)(gamma * tau));
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "rl_gae: launch: %s", cudaGetErrorString(e));
    return 0;
}
