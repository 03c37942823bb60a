// mw_kernels.cu -- World.step of the reference's generic particle world (multiagent/core.py:118-225) for E worlds.
// One thread owns one world: its <= 12 entities live in registers for the step (entity-major planes make every
// global access a coalesced request), the O(NE^2) contact loop is fully unrolled with warp-uniform predicates
// (the configuration is shared by all worlds).  HBM-bound streaming like the FortAttack step:
// 16 B read + 16 B written per movable entity, 8 B read per fixed entity and per action.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mape_world.h"

int fa_internal_fail(int code, const char *fmt, ...);

namespace mw {

constexpr int MAXE = MW_MAX_ENTITIES;

template <typename R> struct V2;
template <> struct V2<float> { typedef float2 T; };
template <> struct V2<double> { typedef double2 T; };

template <typename R> struct Par {
    int E, na, ne;
    R dt, keep, cf, k, xmin, xmax, ymin, ymax;
    R size[MAXE], inv_mass[MAXE], max_speed[MAXE];
    uint32_t collide, movable;      // bit masks
};

// penetration = k * log(1 + exp(-x / k))  (core.py:205,221).  float: with the reference's k = 1e-10 it is max(0, -x) to
// within k ln 2 = 7e-11 (below float resolution); for a margin that float can resolve (the stock MPE 1e-3 commented beside
// it in the reference) the softplus itself is evaluated, so float mode follows the double path / the oracle for any k
__device__ __forceinline__ float pen(float x, float k) {
    if (k <= 1e-6f) return fmaxf(0.0f, -x);
    const float t = -x / k;
    if (t > 30.0f) return -x;
    if (t < -80.0f) return 0.0f;
    return (t > 0.0f ? t + log1pf(expf(-t)) : log1pf(expf(t))) * k;
}
__device__ __forceinline__ double pen(double x, double k) {
    const double t = -x / k;
    if (t > 40.0) return (t + log1p(exp(-t))) * k;
    if (t < -745.0) return 0.0;
    return (t > 0.0 ? t + log1p(exp(-t)) : log1p(exp(t))) * k;
}
__device__ __forceinline__ float sqrt_r(float v) { return sqrtf(v); }
__device__ __forceinline__ double sqrt_r(double v) { return sqrt(v); }

// NE (entities per world) is a template parameter: the loops below unroll exactly, the per-entity state stays in registers
template <typename R, int NE>
__global__ void __launch_bounds__(128) step_kernel(const Par<R> p, typename V2<R>::T *pos, typename V2<R>::T *vel,
                                                   const typename V2<R>::T *u) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.E) return;
    const size_t E = (size_t)p.E;
    R x[NE], y[NE], fx[NE], fy[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        fx[i] = R(0); fy[i] = R(0);
        {
            const typename V2<R>::T q = pos[i * E + e];
            x[i] = q.x; y[i] = q.y;
        }
        if (i < p.na && (p.movable >> i & 1u)) {                       // apply_action_force (core.py:139-145)
            const typename V2<R>::T a = u[i * E + e];
            fx[i] = a.x; fy[i] = a.y;
        }
    }
#pragma unroll
    for (int a = 0; a < NE; ++a) {                                     // apply_environment_force (core.py:148-160,196-210)
#pragma unroll
        for (int b = a + 1; b < NE; ++b) {
            if ((p.collide >> a & 1u) && (p.collide >> b & 1u)) {
                const R dx = x[a] - x[b], dy = y[a] - y[b];
                const R dist = sqrt_r(dx * dx + dy * dy);
                const R g = p.cf * pen(dist - (p.size[a] + p.size[b]), p.k) / dist;     // dist == 0 -> NaN like the reference
                if (p.movable >> a & 1u) { fx[a] += g * dx; fy[a] += g * dy; }
                if (p.movable >> b & 1u) { fx[b] -= g * dx; fy[b] -= g * dy; }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        if (p.movable >> i & 1u) {
            if (i < p.na && (p.collide >> i & 1u)) {                   // apply_wall_collision_force (core.py:163-169,212-225)
                const R s = p.size[i];
                fx[i] += p.cf * (pen(x[i] - s - p.xmin, p.k) - pen(p.xmax - x[i] - s, p.k));
                fy[i] += p.cf * (pen(y[i] - s - p.ymin, p.k) - pen(p.ymax - y[i] - s, p.k));
            }
            typename V2<R>::T v = vel[i * E + e];                      // integrate_state (core.py:172-184)
            v.x = v.x * p.keep + fx[i] * p.inv_mass[i] * p.dt;
            v.y = v.y * p.keep + fy[i] * p.inv_mass[i] * p.dt;
            if (p.max_speed[i] >= R(0)) {
                const R sp = sqrt_r(v.x * v.x + v.y * v.y);
                if (sp > p.max_speed[i]) { v.x = v.x / sp * p.max_speed[i]; v.y = v.y / sp * p.max_speed[i]; }
            }
            vel[i * E + e] = v;
            typename V2<R>::T q;
            q.x = x[i] + v.x * p.dt; q.y = y[i] + v.y * p.dt;
            pos[i * E + e] = q;
        }
    }
}

template <typename R> static cudaError_t launch(const MwConfig &c, void *pos, void *vel, const void *u, cudaStream_t st) {
    Par<R> p;
    p.E = c.n_envs; p.na = c.n_agents; p.ne = c.n_entities;
    p.dt = (R)c.dt; p.keep = (R)(1.0 - c.damping); p.cf = (R)c.contact_force; p.k = (R)c.contact_margin;
    p.xmin = (R)c.wall[0]; p.xmax = (R)c.wall[1]; p.ymin = (R)c.wall[2]; p.ymax = (R)c.wall[3];
    p.collide = p.movable = 0;
    for (int i = 0; i < MAXE; ++i) {
        const bool in = i < c.n_entities;
        p.size[i] = in ? (R)c.size[i] : R(0);
        p.inv_mass[i] = in ? (R)(1.0 / c.mass[i]) : R(0);
        p.max_speed[i] = in ? (R)c.max_speed[i] : R(-1);
        if (in && c.collide[i]) p.collide |= 1u << i;
        if (in && c.movable[i]) p.movable |= 1u << i;
    }
    const int grid = (c.n_envs + 127) / 128;
    typename V2<R>::T *pp = (typename V2<R>::T *)pos, *vv = (typename V2<R>::T *)vel;
    const typename V2<R>::T *uu = (const typename V2<R>::T *)u;
    switch (c.n_entities) {
#define MW_CASE(N) case N: step_kernel<R, N><<<grid, 128, 0, st>>>(p, pp, vv, uu); break;
        MW_CASE(1) MW_CASE(2) MW_CASE(3) MW_CASE(4) MW_CASE(5) MW_CASE(6) MW_CASE(7) MW_CASE(8) MW_CASE(9) MW_CASE(10)
        MW_CASE(11) MW_CASE(12)
#undef MW_CASE
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

}  // namespace mw

extern "C" int mw_step(const MwConfig *cfg, void *d_pos, void *d_vel, const void *d_u, void *stream) {
    if (!cfg || !d_pos || !d_vel || !d_u) return fa_internal_fail(-1, "mw_step: NULL pointer");
    if (cfg->n_envs < 1 || cfg->n_entities < 1 || cfg->n_entities > MW_MAX_ENTITIES || cfg->n_agents < 1 ||
        cfg->n_agents > cfg->n_entities || (cfg->scalar != 0 && cfg->scalar != 1))
        return fa_internal_fail(-1, "mw_step: need 1 <= n_agents <= n_entities <= %d, n_envs >= 1, scalar 0/1", MW_MAX_ENTITIES);
    for (int i = 0; i < cfg->n_entities; ++i)
        if (!(cfg->mass[i] > 0.0)) return fa_internal_fail(-1, "mw_step: entity %d has non-positive mass", i);
    if (((uintptr_t)d_pos | (uintptr_t)d_vel | (uintptr_t)d_u) & 15) return fa_internal_fail(-4, "mw_step: planes must be 16-byte aligned");
    const cudaError_t e = cfg->scalar == 1 ? mw::launch<double>(*cfg, d_pos, d_vel, d_u, (cudaStream_t)stream)
                                           : mw::launch<float>(*cfg, d_pos, d_vel, d_u, (cudaStream_t)stream);
    if (e != cudaSuccess) return fa_internal_fail(-2, "mw_step: launch: %s", cudaGetErrorString(e));
    return 0;
}
