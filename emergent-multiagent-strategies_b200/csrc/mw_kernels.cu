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

// ---- scenario callbacks: observation + reward of every agent after a step, one thread per world -------------------------
// simple_spread (multiagent/scenarios/simple_spread.py:72-101) and simple_tag (simple_tag.py:83-179), evaluated the way
// MultiAgentEnv.step does (environment.py:97-108: per-agent callbacks, then the shared sum when world.collaborative).
template <typename R> struct ScenPar {
    int E, na, ne, scenario, n_adv, obs_stride;
    R size[MAXE];
};

template <typename R> __device__ __forceinline__ R dist2d(R ax, R ay, R bx, R by) {
    const R dx = ax - bx, dy = ay - by;
    return sqrt_r(dx * dx + dy * dy);                                   // np.sqrt(np.sum(np.square(delta_pos)))
}
__device__ __forceinline__ float exp_r(float v) { return expf(v); }
__device__ __forceinline__ double exp_r(double v) { return exp(v); }

template <typename R>
__global__ void __launch_bounds__(128) scenario_kernel(const ScenPar<R> p, const typename V2<R>::T *pos, const typename V2<R>::T *vel,
                                                       R *obs, R *rew) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= p.E) return;
    const size_t E = (size_t)p.E;
    R x[MAXE], y[MAXE];
#pragma unroll
    for (int i = 0; i < MAXE; ++i) {
        if (i < p.ne) { const typename V2<R>::T q = pos[i * E + e]; x[i] = q.x; y[i] = q.y; }
        else { x[i] = R(0); y[i] = R(0); }
    }
    // ---- observation: [vel, pos, landmarks relative, other agents relative, then comm zeros (spread) / velocities of the
    //      other good agents (tag)]
    for (int i = 0; i < p.na; ++i) {
        R *o = obs + ((size_t)i * E + e) * p.obs_stride;
        const typename V2<R>::T v = vel[i * E + e];
        int k = 0;
        o[k++] = v.x; o[k++] = v.y; o[k++] = x[i]; o[k++] = y[i];
        for (int l = p.na; l < p.ne; ++l) { o[k++] = x[l] - x[i]; o[k++] = y[l] - y[i]; }
        for (int j = 0; j < p.na; ++j)
            if (j != i) { o[k++] = x[j] - x[i]; o[k++] = y[j] - y[i]; }
        if (p.scenario == 0) {
            for (int j = 0; j < p.na; ++j)
                if (j != i) { o[k++] = R(0); o[k++] = R(0); }           // other.state.c: agents are silent, dim_c = 2
        } else {
            for (int j = p.n_adv; j < p.na; ++j)
                if (j != i) { const typename V2<R>::T w = vel[j * E + e]; o[k++] = w.x; o[k++] = w.y; }
        }
        for (; k < p.obs_stride; ++k) o[k] = R(0);
    }
    // ---- reward -------------------------------------------------------------------------------------------------------
    if (p.scenario == 0) {
        R cover = R(0);                                                 // -= min over agents of the distance to each landmark
        for (int l = p.na; l < p.ne; ++l) {
            R m = dist2d(x[0], y[0], x[l], y[l]);
            for (int a = 1; a < p.na; ++a) { const R d = dist2d(x[a], y[a], x[l], y[l]); m = d < m ? d : m; }
            cover -= m;
        }
        R total = R(0);
        for (int i = 0; i < p.na; ++i) {
            R r = cover;
            for (int a = 0; a < p.na; ++a)                              // the agent itself counts (dist 0 < 2 size), as in the reference
                if (dist2d(x[a], y[a], x[i], y[i]) < p.size[a] + p.size[i]) r -= R(1);
            total += r;
        }
        for (int i = 0; i < p.na; ++i) rew[i * E + e] = total;          // world.collaborative: everyone gets np.sum(reward_n)
    } else {
        R catches = R(0);
        for (int g = p.n_adv; g < p.na; ++g)
            for (int a = 0; a < p.n_adv; ++a)
                if (dist2d(x[g], y[g], x[a], y[a]) < p.size[g] + p.size[a]) catches += R(10);
        for (int i = 0; i < p.n_adv; ++i) rew[i * E + e] = catches;
        for (int g = p.n_adv; g < p.na; ++g) {
            R r = R(0);
            for (int a = 0; a < p.n_adv; ++a)
                if (dist2d(x[a], y[a], x[g], y[g]) < p.size[a] + p.size[g]) r -= R(10);
            const R c[2] = {x[g] < R(0) ? -x[g] : x[g], y[g] < R(0) ? -y[g] : y[g]};
            for (int d = 0; d < 2; ++d) {                               // bound(): leaving the screen is penalised
                R b = R(0);
                if (c[d] >= R(1)) { const R ex = exp_r(R(2) * c[d] - R(2)); b = ex < R(10) ? ex : R(10); }
                else if (c[d] >= R(0.9)) b = (c[d] - R(0.9)) * R(10);
                r -= b;
            }
            rew[g * E + e] = r;
        }
    }
}

// ---- scenario reset: uniform positions per entity (reset_world of both scenarios), zero velocities --------------------
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

template <typename R>
__global__ void __launch_bounds__(128) reset_kernel(int nE, int ne, int na, R lo_a, R hi_a, R lo_l, R hi_l, uint64_t seed, uint32_t episode,
                                                    const uint8_t *mask, typename V2<R>::T *pos, typename V2<R>::T *vel) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE || (mask != nullptr && !mask[e])) return;
    for (int i = 0; i < ne; ++i) {
        uint32_t c[4] = {(uint32_t)e, (uint32_t)i, episode, 0x4d41u};
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c);
        const double ux = ((double)c[0] + 0.5) * (1.0 / 4294967296.0), uy = ((double)c[1] + 0.5) * (1.0 / 4294967296.0);
        const double lo = i < na ? (double)lo_a : (double)lo_l, hi = i < na ? (double)hi_a : (double)hi_l;
        typename V2<R>::T q, z;
        q.x = (R)(lo + (hi - lo) * ux); q.y = (R)(lo + (hi - lo) * uy);
        z.x = R(0); z.y = R(0);
        pos[(size_t)i * nE + e] = q;
        vel[(size_t)i * nE + e] = z;
    }
}

}  // namespace mw

extern "C" int mw_scenario_obs_dim(int scenario, int n_agents, int n_entities, int n_adversaries) {
    const int L = n_entities - n_agents;
    if (scenario == MW_SIMPLE_SPREAD) return 4 + 2 * L + 4 * (n_agents - 1);
    if (scenario == MW_SIMPLE_TAG) return 4 + 2 * L + 2 * (n_agents - 1) + 2 * (n_agents - n_adversaries);   // widest row (an adversary's)
    return -1;
}

extern "C" int mw_scenario_callbacks(const MwConfig *cfg, int scenario, int n_adversaries, const void *d_pos, const void *d_vel,
                                     void *d_obs, int obs_stride, void *d_reward, void *stream) {
    if (!cfg || !d_pos || !d_vel || !d_obs || !d_reward) return fa_internal_fail(-1, "mw_scenario_callbacks: NULL pointer");
    if (cfg->n_envs < 1 || cfg->n_entities < 1 || cfg->n_entities > MW_MAX_ENTITIES || cfg->n_agents < 1 ||
        cfg->n_agents > cfg->n_entities || (cfg->scalar != 0 && cfg->scalar != 1))
        return fa_internal_fail(-1, "mw_scenario_callbacks: bad configuration");
    if (scenario != MW_SIMPLE_SPREAD && scenario != MW_SIMPLE_TAG)
        return fa_internal_fail(-1, "mw_scenario_callbacks: scenario must be MW_SIMPLE_SPREAD or MW_SIMPLE_TAG");
    if (scenario == MW_SIMPLE_TAG && (n_adversaries < 1 || n_adversaries >= cfg->n_agents))
        return fa_internal_fail(-1, "mw_scenario_callbacks: simple_tag needs 1 <= n_adversaries < n_agents");
    if (obs_stride < mw_scenario_obs_dim(scenario, cfg->n_agents, cfg->n_entities, n_adversaries))
        return fa_internal_fail(-1, "mw_scenario_callbacks: obs_stride %d is smaller than the observation", obs_stride);
    const int grid = (cfg->n_envs + 127) / 128;
    if (cfg->scalar == 1) {
        mw::ScenPar<double> p = {cfg->n_envs, cfg->n_agents, cfg->n_entities, scenario, n_adversaries, obs_stride, {}};
        for (int i = 0; i < mw::MAXE; ++i) p.size[i] = i < cfg->n_entities ? cfg->size[i] : 0.0;
        mw::scenario_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(p, (const double2 *)d_pos, (const double2 *)d_vel, (double *)d_obs, (double *)d_reward);
    } else {
        mw::ScenPar<float> p = {cfg->n_envs, cfg->n_agents, cfg->n_entities, scenario, n_adversaries, obs_stride, {}};
        for (int i = 0; i < mw::MAXE; ++i) p.size[i] = i < cfg->n_entities ? (float)cfg->size[i] : 0.0f;
        mw::scenario_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(p, (const float2 *)d_pos, (const float2 *)d_vel, (float *)d_obs, (float *)d_reward);
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "mw_scenario_callbacks: launch: %s", cudaGetErrorString(e));
    return 0;
}

extern "C" int mw_scenario_reset(const MwConfig *cfg, double lo_agents, double hi_agents, double lo_landmarks, double hi_landmarks,
                                 uint64_t seed, uint32_t episode, const uint8_t *d_mask, void *d_pos, void *d_vel, void *stream) {
    if (!cfg || !d_pos || !d_vel) return fa_internal_fail(-1, "mw_scenario_reset: NULL pointer");
    if (cfg->n_envs < 1 || cfg->n_entities < 1 || cfg->n_entities > MW_MAX_ENTITIES || cfg->n_agents > cfg->n_entities ||
        (cfg->scalar != 0 && cfg->scalar != 1))
        return fa_internal_fail(-1, "mw_scenario_reset: bad configuration");
    const int grid = (cfg->n_envs + 127) / 128;
    if (cfg->scalar == 1)
        mw::reset_kernel<double><<<grid, 128, 0, (cudaStream_t)stream>>>(cfg->n_envs, cfg->n_entities, cfg->n_agents, lo_agents, hi_agents,
                                                                         lo_landmarks, hi_landmarks, seed, episode, d_mask, (double2 *)d_pos, (double2 *)d_vel);
    else
        mw::reset_kernel<float><<<grid, 128, 0, (cudaStream_t)stream>>>(cfg->n_envs, cfg->n_entities, cfg->n_agents, (float)lo_agents, (float)hi_agents,
                                                                        (float)lo_landmarks, (float)hi_landmarks, seed, episode, d_mask, (float2 *)d_pos, (float2 *)d_vel);
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fa_internal_fail(-2, "mw_scenario_reset: launch: %s", cudaGetErrorString(e));
    return 0;
}

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
