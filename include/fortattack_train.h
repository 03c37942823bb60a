/*
 * fortattack_train.h -- C ABI of the dense products and the optimizer step of the PPO update
 * (libfortattack_b200.so), hand-written for sm_100a on the tcgen05 tensor cores.
 *
 * What they replace in the reference (all of it runs there as torch/cuBLAS fp32 library calls):
 *     MPNN._fwd / evaluate_actions and its autograd backward    mpnn.py:117-205, 289-327, 409-437
 *         nn.Linear / torch.matmul / torch.mm forward                 -> tg_linear      (y = act(x W^T + b))
 *         the input gradient of those layers  (dx = dy W)             -> tg_linear      (with the transposed pack)
 *         the weight gradient of those layers (dW = dy^T x)           -> tg_wgrad
 *     nn.utils.clip_grad_norm_ + optim.Adam.step                      rlcore/algo/ppo.py:114,189-192
 *                                                                     -> tg_adam_step
 *
 * Arithmetic: "fp32-grade" on 16-bit tensor-core operands.  Every fp32 operand is split into 16-bit terms and the
 * product is the sum of the cross terms that matter, accumulated in fp32 in tensor memory:
 *   tg_linear  x = (x_hi + x_lo) / s_row, W = (W_hi + W_lo) / s_W, fp16 terms, s = power of two that puts the largest
 *              |value| of the row / of the tensor at 2^14 (nothing lands in fp16's subnormal range that matters);
 *              x W^T ~ x_hi W_hi + x_hi W_lo + x_lo W_hi: three MMAs per product, relative error ~2^-21
 *   tg_wgrad   the reduction runs over the ROWS, so a per-row scale cannot be factored out; operands are split into
 *              three bf16 terms (8-bit exponent: no scaling needed) and six cross products are accumulated, ~2^-23
 * (single-pass tf32 / fp16 / bf16 operands miss this repo's gradient gate by 2-3 orders of magnitude:
 * profiles/r1j_training_precision_study.txt).
 *
 * Conventions as in fortattack.h: plain C types, caller-owned device memory, explicit stream, 0 / negative error code
 * with the message in fa_last_error().  All matrices are row-major fp32 with a row stride ("ld", in floats).
 */
#ifndef FORTATTACK_TRAIN_B200_H
#define FORTATTACK_TRAIN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Bytes of the packed form of a [N][K] weight (N, K as given to tg_pack_weight; padding included). */
size_t tg_packed_bytes(int N, int K);

/* Pack the B operand of y = x B^T:  B[n][k] = transposed ? d_w[k * ld + n] : d_w[n * ld + k]  (n < N, k < K) into
 * fp16 hi / lo planes in the tensor cores' K-major core-matrix order, scaled by a power of two; the inverse scale is
 * stored with the planes.  N <= 256, K <= 256.  (One small launch; weights change every optimizer step.) */
int tg_pack_weight(const float *d_w, int N, int K, int ld, int transposed, void *d_packed, void *stream);

/* out[r][n] = act( sum_k x[r][k] B[n][k] + bias[n] ) (+ out[r][n] if accumulate)   r < rows, n < N, k < K
 *   d_x      float [rows][K], row stride ldx            d_packed  tg_pack_weight(B, N, K)
 *   d_bias   float [N] or NULL                          relu      apply max(0, .) after the bias
 *   d_out    float [rows][N], row stride ldo
 * One persistent CTA per SM: loader warps convert 128-row tiles of x to fp16 hi/lo in shared memory, one thread issues
 * tcgen05.mma against the resident weight planes, epilogue warps drain the accumulators from tensor memory. */
int tg_linear(const float *d_x, int ldx, long long rows, int K, const void *d_packed, int N, const float *d_bias, int relu,
              int accumulate, float *d_out, int ldo, uint32_t *d_status, void *stream);

/* The same with the accumulated term read from its own matrix:  out = act(x B^T + bias) + res  (d_res float [rows][N], row
 * stride ldr; NULL = no accumulated term; d_res == d_out is tg_linear's accumulate form).  dh = dh_direct + dg Mqk^T without a
 * copy of dh_direct (mpnn.py's backward through the encoder / message rounds). */
int tg_linear_res(const float *d_x, int ldx, long long rows, int K, const void *d_packed, int N, const float *d_bias, int relu,
                  const float *d_res, int ldr, float *d_out, int ldo, uint32_t *d_status, void *stream);

/* Bytes of scratch tg_wgrad needs for an [a][b] result. */
size_t tg_wgrad_scratch_bytes(int a, int b);

/* d_dw[i][j] (+)= sum_r x[r][i] y[r][j]      i < a <= 256, j < b <= 256, r < rows      (the weight gradient dW = dy^T x of
 * a dense layer: x = dy, y = layer input; mpnn.py's backward)
 *   d_x float [rows][a] (ldx), d_y float [rows][b] (ldy), d_dw float [a][b] (lddw); d_scratch: tg_wgrad_scratch_bytes
 * Every CTA accumulates its share of the rows in tensor memory and writes one partial [a][b] block; a second small kernel
 * adds the partials in a fixed order (bit-reproducible, no atomics). */
int tg_wgrad(const float *d_x, int ldx, int a, const float *d_y, int ldy, int b, long long rows, float *d_dw, int lddw,
             int accumulate, void *d_scratch, uint32_t *d_status, void *stream);

/* One optimizer step for up to TG_MAX_TENSORS parameter tensors (rlcore/algo/ppo.py:189-192 after the backward):
 *   total = sqrt(sum_t |g_t|^2);  c = min(1, max_norm / (total + 1e-6))          nn.utils.clip_grad_norm_
 *   g *= c;  m = b1 m + (1 - b1) g;  v = b2 v + (1 - b2) g^2;  step += 1
 *   p -= lr / (1 - b1^step) * m / (sqrt(v) / sqrt(1 - b2^step) + eps)            torch.optim.Adam (no amsgrad, no decay)
 * Two launches (norm partials, update); the step counter lives on the device (d_step int64 [1]) so the call can be
 * captured in a CUDA graph.  A NULL gradient pointer skips that tensor (parameters that received no gradient).
 * d_scratch: TG_ADAM_SCRATCH_FLOATS floats.  max_norm <= 0 disables clipping.  d_total_norm (optional) receives total.
 * d_grad_scale (optional): device scalar every gradient is multiplied by first (in place) -- the loss normaliser /
 * world size of a multi-rank step whose gradients were all-reduced as un-normalised sums. */
#define TG_MAX_TENSORS 32
#define TG_ADAM_SCRATCH_FLOATS 256
typedef struct TgTensor {
    float *p, *g, *m, *v;
    long long numel;
} TgTensor;
int tg_adam_step(const TgTensor *tensors, int n_tensors, float lr, float beta1, float beta2, float eps, float max_norm,
                 const float *d_grad_scale, long long *d_step, float *d_scratch, float *d_total_norm, void *stream);

/* Static facts for reports: registers / block / dynamic shared memory of the two GEMM kernels. */
int tg_kernel_info(int which, int32_t *regs, int32_t *block, int32_t *smem);

/* tg_linear's epilogue leaves through the TMA engine (one cp.async.bulk.tensor store per 32 x 16 block; an accumulated term
 * arrives by tensor-map loads of the same blocks) whenever N >= 16 and the rows of the output (and of the added matrix) are
 * 16-byte multiples apart; tg_debug_tma_out(0) keeps it on plain loads / stores.  Both epilogues compute the same values in
 * the same order (bit-equal, test). */
int tg_debug_tma_out(int on);

/* tg_wgrad stages its operands in blocks of 64 rows (ring of 4), or of 32 rows (ring of 8) when an operand is wider than 128
 * columns (three blocks per row step).  tg_debug_wgrad_rows(32 | 64) forces one height, 0 restores the choice by shape. */
int tg_debug_wgrad_rows(int rows);
/* tg_wgrad has the same two loader forms as tg_linear: STAGED (both operands in whole 8-column chunks with 16-byte aligned rows:
 * a producer warp fetches the raw fp32 rows of every operand block with one tensor-map load, the loader warps convert from
 * shared memory; operand ring of 3 / 6 blocks + a 64 KB raw ring) and the register form (operand ring of 4 / 8 blocks).  Same
 * arithmetic and the same partition of the rows: bit-equal results (test).  tg_debug_wgrad_staged(0) keeps the register form. */
int tg_debug_wgrad_staged(int on);

/* tg_linear has two loader forms with identical arithmetic: the STAGED form (K = 64 or 128, 16-byte aligned rows: a producer
 * warp streams raw fp32 rows into a shared-memory ring with bulk copies and the loader warps convert from there) and the
 * register form (every other shape).  tg_debug_staged(0) keeps every call on the register form (measurements, tests of
 * both forms); tg_debug_staged(1) restores the default. */
int tg_debug_staged(int on);

/* Debug: override the descriptor fields of the MN-major operands of tg_wgrad (lbo, sbo in bytes; 0 = built-in). */
int tg_debug_wgrad_desc(uint32_t lbo, uint32_t sbo);

#ifdef __cplusplus
}
#endif
#endif
