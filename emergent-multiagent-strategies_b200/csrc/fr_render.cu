// fr_render.cu -- device-state rasteriser: the scene of FortAttackGlobalEnv.render (gym_fortattack/fortattack.py:368-596)
// for selected environments of a batch, from the observation planes.  C ABI and the paint order: include/fortattack_render.h.
//
// One thread per pixel, one block = 16 x 16 pixels of one image.  The agents' geometry (body / head centres, laser triangle,
// halo radius, colour) is prepared once per block by the first A threads -- trigonometry in double, rounded once to float --
// and every per-pixel predicate and blend is a single correctly rounded float operation (__fmul_rn / __fadd_rn / __fsub_rn,
// no FMA contraction), so the numpy float32 restatement the tests use reproduces the image bit for bit.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/fortattack_render.h"

int fa_internal_fail(int code, const char *fmt, ...);

namespace fr {

constexpr int MAX_A = 10;
constexpr float SIZE = 0.05f;                  // Entity.size (core.py:32)
constexpr double HEAD_SHIFT = 0.8 * 0.05;      // fortattack.py:505
constexpr double SHOOT_RAD = 0.8;              // core.py:100
constexpr double HALF_WIN = 0.39269908169872414;   // shootWin / 2 = pi / 8 (core.py:101, 373-382)

struct Agent {
    float cx, cy, hx, hy;         // body and head centres
    float lx[3], ly[3];           // laser triangle
    float halo_r2;                // < 0: none
    float r, g, b;
    int alive, laser;
};

__device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }

__device__ __forceinline__ bool in_disc(float wx, float wy, float cx, float cy, float r2) {
    const float dx = sub(wx, cx), dy = sub(wy, cy);
    return add(mul(dx, dx), mul(dy, dy)) <= r2;
}
__device__ __forceinline__ float edge(float ax, float ay, float bx, float by, float px, float py) {
    return sub(mul(sub(bx, ax), sub(py, ay)), mul(sub(by, ay), sub(px, ax)));
}
__device__ __forceinline__ void blend(float (&c)[3], float r, float g, float b, float a) {
    const float om = sub(1.0f, a);
    c[0] = add(mul(c[0], om), mul(r, a));
    c[1] = add(mul(c[1], om), mul(g, a));
    c[2] = add(mul(c[2], om), mul(b, a));
}
__device__ __forceinline__ void paint(float (&c)[3], float r, float g, float b) { c[0] = r; c[1] = g; c[2] = b; }

__device__ __forceinline__ void draw_agent(float (&c)[3], const Agent &a, float wx, float wy, float scale) {
    const float r = mul(a.r, scale), g = mul(a.g, scale), b = mul(a.b, scale);
    const float head_r2 = mul(mul(0.5f, SIZE), mul(0.5f, SIZE)), body_r2 = mul(SIZE, SIZE);
    if (in_disc(wx, wy, a.hx, a.hy, head_r2)) paint(c, r, g, b);
    if (a.laser) {
        const float e0 = edge(a.lx[0], a.ly[0], a.lx[1], a.ly[1], wx, wy), e1 = edge(a.lx[1], a.ly[1], a.lx[2], a.ly[2], wx, wy),
                    e2 = edge(a.lx[2], a.ly[2], a.lx[0], a.ly[0], wx, wy);
        if ((e0 >= 0.f && e1 >= 0.f && e2 >= 0.f) || (e0 <= 0.f && e1 <= 0.f && e2 <= 0.f)) blend(c, r, g, b, 0.3f);
    }
    if (in_disc(wx, wy, a.cx, a.cy, body_r2)) paint(c, r, g, b);
}

__global__ void __launch_bounds__(256) render_kernel(const FrConfig cfg, const float *__restrict__ obs, const int32_t *__restrict__ actions,
                                                     const float *__restrict__ halo, const int32_t *__restrict__ env_ids,
                                                     uint8_t *__restrict__ rgb) {
    __shared__ Agent ag[MAX_A];
    const int A = cfg.n_guards + cfg.n_attackers, E = cfg.n_envs;
    const int img = blockIdx.z, t = threadIdx.y * blockDim.x + threadIdx.x;
    const int e = env_ids[img];
    if (t < A) {
        Agent a;
        const float *o = obs + ((size_t)t * E + e) * 6;
        const float x = o[1], y = o[2];
        const double ang = (double)o[3], cs = cos(ang), sn = sin(ang);
        a.alive = o[0] != 0.f;
        a.cx = x; a.cy = y;
        a.hx = (float)((double)x + HEAD_SHIFT * cs);
        a.hy = (float)((double)y + HEAD_SHIFT * sn);
        const double p1x = (double)x + (double)SIZE * cs, p1y = (double)y + (double)SIZE * sn;     // get_tri_pts_arr, core.py:373-382
        a.lx[0] = (float)p1x; a.ly[0] = (float)p1y;
        a.lx[1] = (float)(p1x + SHOOT_RAD * cos(ang + HALF_WIN)); a.ly[1] = (float)(p1y + SHOOT_RAD * sin(ang + HALF_WIN));
        a.lx[2] = (float)(p1x + SHOOT_RAD * cos(ang - HALF_WIN)); a.ly[2] = (float)(p1y + SHOOT_RAD * sin(ang - HALF_WIN));
        a.laser = actions != nullptr && actions[(size_t)t * E + e] == 7;
        a.halo_r2 = -1.0f;
        if (halo != nullptr) {
            const float w = halo[(size_t)t * E + e];
            if (w >= 0.f) {
                const float r = mul(SIZE, add(1.0f, w));
                a.halo_r2 = mul(r, r);
            }
        }
        const bool guard = t < cfg.n_guards;                       // fortattack_env_v1.py:57
        a.r = guard ? 0.f : 1.f; a.g = guard ? 1.f : 0.f; a.b = 0.f;
        ag[t] = a;
    }
    __syncthreads();
    const int px = blockIdx.x * blockDim.x + threadIdx.x, py = blockIdx.y * blockDim.y + threadIdx.y;
    if (px >= cfg.width || py >= cfg.height) return;
    const float sx = __fdiv_rn(2.0f, (float)cfg.width), sy = __fdiv_rn(2.0f, (float)cfg.height);
    const float wx = sub(mul(add((float)px, 0.5f), sx), 1.0f), wy = sub(1.0f, mul(add((float)py, 0.5f), sy));
    float c[3] = {1.f, 1.f, 1.f};                                                       // glClearColor(1,1,1,1), rendering.py:90
    if (wx >= -1.0f && wx <= 1.0f && wy >= -0.8f && wy <= 0.8f) paint(c, 0.f, 0.f, 0.f);   // world rectangle
    if (in_disc(wx, wy, 0.0f, 0.8f, mul(0.15f, 0.15f))) paint(c, 0.f, 1.f, 1.f);        // fort
    for (int i = 0; i < A; ++i)                                                         // attention halos
        if (ag[i].halo_r2 >= 0.f && (ag[i].alive || cfg.draw_dead) && in_disc(wx, wy, ag[i].cx, ag[i].cy, ag[i].halo_r2))
            blend(c, 1.f, 1.f, 0.f, ag[i].alive ? 0.9f : 0.3f);
    if (cfg.draw_dead)
        for (int i = 0; i < A; ++i)
            if (!ag[i].alive) {                                                         // no laser for the dead (core.py:268 loops over alive)
                Agent d = ag[i];
                d.laser = 0;
                draw_agent(c, d, wx, wy, 0.5f);
            }
    for (int i = 0; i < A; ++i)
        if (ag[i].alive) draw_agent(c, ag[i], wx, wy, 1.0f);
    if (wy > 0.8f || wy < -0.8f) paint(c, 0.5f, 0.5f, 0.5f);                            // grey strips
    uint8_t *out = rgb + (((size_t)img * cfg.height + py) * cfg.width + px) * 3;
#pragma unroll
    for (int k = 0; k < 3; ++k) out[k] = (uint8_t)__float2int_rd(add(mul(c[k], 255.0f), 0.5f));
}

}  // namespace fr

extern "C" int fr_render(const FrConfig *cfg, const float *d_obs, const int32_t *d_actions, const float *d_halo,
                         const int32_t *d_env_ids, int n_img, uint8_t *d_rgb, void *stream) {
    if (!cfg || !d_obs || !d_env_ids || !d_rgb) return fa_internal_fail(-1, "fr_render: NULL pointer");
    const int A = cfg->n_guards + cfg->n_attackers;
    if (cfg->n_envs < 1 || cfg->n_guards < 1 || cfg->n_attackers < 1 || cfg->n_guards > 5 || cfg->n_attackers > 5 || A > fr::MAX_A)
        return fa_internal_fail(-1, "fr_render: need n_envs >= 1 and 1..5 guards / attackers (got %d envs, %dv%d)", cfg->n_envs,
                                cfg->n_guards, cfg->n_attackers);
    if (cfg->width < 1 || cfg->height < 1 || cfg->width > 8192 || cfg->height > 8192 || n_img < 1 || n_img > 65535)
        return fa_internal_fail(-1, "fr_render: image size %d x %d / count %d out of range", cfg->width, cfg->height, n_img);
    const dim3 block(16, 16), grid((cfg->width + 15) / 16, (cfg->height + 15) / 16, n_img);
    fr::render_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*cfg, d_obs, d_actions, d_halo, d_env_ids, d_rgb);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "fr_render: launch: %s", cudaGetErrorString(e));
    return 0;
}
