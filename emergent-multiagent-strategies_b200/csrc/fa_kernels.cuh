// fa_kernels.cuh -- the fused FortAttack step kernel for sm_100a (B200).
//
// One thread owns one environment for the whole step: the state of its A agents is loaded from
// agent-major SoA planes ([A][E], so that the 32 lanes of a warp read 32 consecutive elements of
// one plane: every global access is a fully coalesced 128/256/512-byte request), lives in registers
// while the reference's whole env.step() pipeline runs on it, and is written back in place:
//
//   decode        gym_fortattack/fortattack.py:235-302    (_set_action)
//   laser         gym_fortattack/core.py:254-302,373-390  (apply_laser_effect / get_tri_pts_arr / laser_hit)
//   action force  gym_fortattack/core.py:221-228
//   contact       gym_fortattack/core.py:231-243,440-456
//   wall          gym_fortattack/core.py:246-252,459-472
//   integrate     gym_fortattack/core.py:305-338
//   observation   gym_fortattack/envs/fortattack_env_v1.py:191-238
//   reward        gym_fortattack/envs/fortattack_env_v1.py:87-188
//   done          gym_fortattack/fortattack.py:202-225, :171
//   reset         gym_fortattack/envs/fortattack_env_v1.py:47-75   (fused auto-reset, Philox4x32-10)
//
// The O(A^2) contact and laser tests are therefore register-to-register (fully unrolled over the
// compile-time team sizes), with no inter-thread exchange at all.  Observations ([A][E][6], 24-byte
// rows) are the one output whose natural per-thread store is strided, so they are transposed through
// a per-warp shared-memory stage and leave as 16-byte vector stores covering whole 32-byte sectors.
//
// The same template is instantiated for float (production) and double (parity mode).  No tensor-core
// work exists on this path: it is HBM-bound streaming (88 B per agent-step + 12 B per
// env-step against ~150 instructions per agent-step).
#pragma once

#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace fa {

// ------------------------------------------------------------------------------------------------
// flags word (one uint32 per agent)
//   bit 0 alive, 1 justDied, 2 hit, 3 wasHit            core.py:90-94,103
//   bits 8-11 numHit, 12-15 numWasHit (saturating at 15; each is bounded by the opposing team size)
//   bits 16-31 turn count of the heading (float mode only, see AngOps<float>)
enum : uint32_t {
    F_ALIVE = 1u, F_JD = 2u, F_HIT = 4u, F_WASHIT = 8u,
    F_NHIT_SHIFT = 8, F_NWAS_SHIFT = 12, F_CNT_MASK = 15u, F_WRAP_SHIFT = 16
};

template <typename R> struct VecT;
template <> struct VecT<float> {
    typedef float4 T4;
    typedef float2 T2;
};
struct __align__(32) Double4 { double x, y, z, w; };
template <> struct VecT<double> {
    typedef Double4 T4;
    typedef double2 T2;
};

template <typename R> struct StateView {
    typename VecT<R>::T4 *pv;   // [A][E]  x, y, vx, vy
    typename VecT<R>::T2 *ap;   // [A][E]  ang (float: reduced to [0,2pi)), prevDist (NaN = None)
    uint32_t *fl;               // [A][E]  flags word
    int32_t *tstep;             // [E]     world.time_step
    uint32_t *episode;          // [E]     resets so far
};

template <typename R> struct StepParams {
    StateView<R> st;
    const int32_t *act;   // [T][A][E]
    R *obs;               // [T][A][E][6]  (may be null)
    R *rew;               // [T][A][E]     (may be null)
    uint8_t *done;        // [T][E]        (may be null)
    uint8_t *result;      // [T][E]        (may be null)
    int E, T, max_steps, auto_reset, obs_vec_ok;
    uint8_t *alive_end;   // [T][E] (may be null): alive guards | alive attackers << 4 at the END of the step, before any reset
    int pdl;              // launched with programmatic stream serialization: the state is read after griddepcontrol.wait
    uint64_t seed, env_id0;
    // optional rollout bookkeeping, fused into the step (fa_set_rollout_outputs; train_fortattack.py:53,97-104, storage.py:41):
    float *mask_next;     // [T][A][E]  done ? alive flag of the (reset) observation : alive flag before the step
    uint8_t *end_next;    // [T][E]     done
    float *ep_rew;        // [A][E]     += reward * alive flag before the step
};

// the three bookkeeping outputs of one (agent, env) pair; alive0 = alive before the step, alive2 = alive in the new observation
template <typename R>
__device__ __forceinline__ void rollout_outputs(const StepParams<R> &p, size_t t, int A, size_t E, int i, int e, bool alive0,
                                                bool alive2, bool dn, R rew) {
    const size_t k = (size_t)i * E + e;
    if (p.mask_next != nullptr) p.mask_next[t * A * E + k] = (dn ? alive2 : alive0) ? 1.0f : 0.0f;
    if (p.ep_rew != nullptr && alive0) p.ep_rew[k] += (float)rew;
    if (p.end_next != nullptr && i == 0) p.end_next[t * E + e] = dn ? 1 : 0;
}

// ------------------------------------------------------------------------------------------------
// scalar helpers (float: SFU approximations are ~1e-7 relative, far inside the 1e-5 parity budget)
__device__ __forceinline__ float rsqrt_t(float v) { return rsqrtf(v); }
__device__ __forceinline__ double rsqrt_t(double v) { return 1.0 / sqrt(v); }
// rsqrtf pinned where it is written (the select-style code below wants it issued unconditionally, ahead of its use)
__device__ __forceinline__ float rsqrt_pinned(float v) { float r; asm volatile("rsqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v)); return r; }
__device__ __forceinline__ double rsqrt_pinned(double v) { return 1.0 / sqrt(v); }
__device__ __forceinline__ float sqrt_t(float v) { return sqrtf(v); }
__device__ __forceinline__ double sqrt_t(double v) { return sqrt(v); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }

// Penetration depth for overlap t (= dist_min - dist, or -wall_distance): the reference's softplus
// k*log(1+exp(t/k)) with k = contact_margin = 1e-10 (core.py:452,468).  It differs from max(0,t) by at
// most k*ln2 = 6.9e-11 -- invisible in float, but agents resting exactly on a wall (t == 0) do see it
// in float64, so double mode evaluates the softplus itself (np.logaddexp(0, t/k) * k).
__device__ __forceinline__ float penetration(float t) { return fmaxf(0.0f, t); }
__device__ __forceinline__ double penetration(double t) {
    const double k = 1e-10, x = t / k;
    if (x > 40.0) return (x + log1p(exp(-x))) * k;
    if (x < -745.0) return 0.0;
    return (x > 0.0 ? x + log1p(exp(-x)) : log1p(exp(x))) * k;
}
// squared centre distance below which a pair can exchange a contact force that is not < 1e-16
template <typename R> struct ContactGate;
template <> struct ContactGate<float> { static constexpr float R2 = 0.1f * 0.1f; };
template <> struct ContactGate<double> { static constexpr double R2 = (0.1 + 1e-7) * (0.1 + 1e-7); };

// ------------------------------------------------------------------------------------------------
// scenario constants: core.py:32,100-101,115-128 ; fortattack_env_v1.py:16-17,33-35
template <typename R> struct K {
    static constexpr R SIZE = R(0.05), SHOOT_RAD = R(0.8), DT = R(0.1), DAMP = R(0.75);
    static constexpr R CONTACT_FORCE = R(100), DIST_MIN = R(0.1), DIST_MIN2 = R(0.1) * R(0.1);
    static constexpr R WALL_X = R(0.95), WALL_Y = R(0.75);        // wall - size
    static constexpr R FORT_DIM = R(0.15), DOOR_Y = R(0.8), GUARD_RING = R(0.3);
    static constexpr R ACCEL = R(3), MAX_SPEED = R(3), MAX_ROT = R(0.17);
    static constexpr R PI = R(3.14159265358979323846), TWO_PI = R(6.28318530717958647692);
    // laser triangle in the shooter's frame: apex at p1, half-angle pi/8, side 0.8 (core.py:373-382)
    static constexpr R INV_LC8 = R(1.0 / (0.8 * 0.92387953251128675613));   // 1/(L cos(pi/8))
    static constexpr R INV_LS8 = R(1.0 / (0.8 * 0.38268343236508977173));   // 1/(L sin(pi/8))
};

// ------------------------------------------------------------------------------------------------
// heading representation
//  double: the raw accumulated angle, exactly as the reference keeps it (core.py:336).
//  float : the reference adds (u mod 2pi) >= 0 every step, so headings grow past 100 rad within an
//          episode, where a float ulp is 1.5e-5.  Float mode therefore keeps ang mod 2pi (ulp 4.8e-7)
//          plus an integer turn count in the flags word; trig uses the reduced angle, and the
//          observation is reassembled as reduced + turns*2pi with a split constant so that it is the
//          float nearest to the reference's double value.
template <typename R> struct AngOps;

template <> struct AngOps<double> {
    static __device__ __forceinline__ void advance(double &a, uint32_t &, int act) {
        // Python: p_ang += u[2] % (2*pi): +0.17 -> 0.17, -0.17 -> 2pi-0.17 (core.py:336)
        if (act == 5) a += 0.17;
        if (act == 6) a += fmod(-0.17, 6.283185307179586) + 6.283185307179586;
    }
    static __device__ __forceinline__ void sincos_heading(double a, double &s, double &c) { sincos(a, &s, &c); }
    static __device__ __forceinline__ double full(double a, uint32_t) { return a; }
    static __device__ __forceinline__ void set_reset(double &a, uint32_t &, bool attacker) {
        a = attacker ? 1.5707963267948966 : 4.71238898038469;   // pi/2 , 3pi/2 (v1:60)
    }
};

template <> struct AngOps<float> {
    // 2pi = HI + LO with HI exact in 9 bits: turns*HI is exact for turns < 2^15
    static constexpr float HI = 6.28125f, LO = 1.9353071795864769e-3f;
    static __device__ __forceinline__ void advance(float &a, uint32_t &fl, int act) {
        if (act == 5) {
            a += 0.17f;
            if (a >= 6.2831853f) { a = (a - HI) - LO; fl += 1u << F_WRAP_SHIFT; }
        }
        if (act == 6) {   // + (2pi - 0.17): one more turn unless the reduced angle borrows it back
            a -= 0.17f;
            if (a < 0.0f) a = (a + HI) + LO; else fl += 1u << F_WRAP_SHIFT;
        }
    }
    static __device__ __forceinline__ void sincos_heading(float a, float &s, float &c) {
        // a in [0,2pi): evaluate at a-pi in [-pi,pi) where the SFU path is specified to 2^-21.4 abs
        float ss, cc;
        __sincosf((a - 3.140625f) - 9.67653589793e-4f, &ss, &cc);
        s = -ss; c = -cc;
    }
    static __device__ __forceinline__ float full(float a, uint32_t fl) {
        float w = (float)(fl >> F_WRAP_SHIFT);
        return fmaf(w, HI, fmaf(w, LO, a));
    }
    static __device__ __forceinline__ void set_reset(float &a, uint32_t &fl, bool attacker) {
        a = attacker ? 1.57079633f : 4.71238898f;
        fl &= (1u << F_WRAP_SHIFT) - 1u;
    }
};

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. 2011), same keying as the oracle: key = seed, counter = (env id lo,
// env id hi, episode, agent pair); lanes 0,1 -> agent 2*pair (x,y), lanes 2,3 -> agent 2*pair+1.
__device__ __forceinline__ void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c[0]), l0 = 0xD2511F53u * c[0];
        uint32_t h1 = __umulhi(0xCD9E8D57u, c[2]), l1 = 0xCD9E8D57u * c[2];
        uint32_t n0 = h1 ^ c[1] ^ k0, n2 = h0 ^ c[3] ^ k1;
        c[0] = n0; c[1] = l1; c[2] = n2; c[3] = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__device__ __forceinline__ double u01(uint32_t r) {
    return __dmul_rn(__dadd_rn((double)r, 0.5), 1.0 / 4294967296.0);
}

// Programmatic dependent launch (sm_90+): a step launched with the attribute may start while the previous step of
// the same stream is still running; everything it does before pdl_wait() (index arithmetic, fetching this step's
// actions, which no step kernel writes) overlaps the predecessor's tail and the launch latency.  pdl_wait() returns
// once the predecessor grid has completed and its writes are visible.  Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
template <int A, typename R> struct Env {
    R x[A], y[A], vx[A], vy[A], ang[A], pd[A];
    uint32_t fl[A];
    int32_t t;
};

// reset_world (fortattack_env_v1.py:47-75): alive, zero velocity, team heading, uniform spawn boxes;
// hit flags and counters cleared; prevDist and justDied are NOT touched.
// Products/sums are individually rounded (no FMA contraction) so that double mode is bit-equal to
// the oracle and float mode is its correctly rounded image.
template <int NG, int NA, typename R>
__device__ __forceinline__ void reset_env(Env<NG + NA, R> &s, uint64_t seed, uint64_t env_id, uint32_t episode) {
    constexpr int A = NG + NA;
    s.t = 0;
#pragma unroll
    for (int pair = 0; pair * 2 < A; ++pair) {
        uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, (uint32_t)pair};
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int i = pair * 2 + h;
            if (i < A) {
                const bool attacker = i >= NG;
                double ux = u01(c[2 * h]), uy = u01(c[2 * h + 1]);
                double px, py;
                if (attacker) {   // x ~ U(-1,1), y ~ U(-0.8,-0.64)   (v1:66)
                    px = __dadd_rn(-1.0, __dmul_rn(1.0 - (-1.0), ux));
                    py = __dadd_rn(-0.8, __dmul_rn(0.8 * -0.8 - (-0.8), uy));
                } else {          // x ~ U(-0.06,0.06), y ~ U(0.64,0.8) (v1:70)
                    const double lo = -0.8 * 0.15 / 2, hi = 0.8 * 0.15 / 2;
                    px = __dadd_rn(lo, __dmul_rn(hi - lo, ux));
                    py = __dadd_rn(0.8 * 0.8, __dmul_rn(0.8 - 0.8 * 0.8, uy));
                }
                s.x[i] = (R)px; s.y[i] = (R)py;
                s.vx[i] = R(0); s.vy[i] = R(0);
                uint32_t f = s.fl[i];
                f = (f & (F_JD | (~0u << F_WRAP_SHIFT))) | F_ALIVE;   // keep justDied; clear hit/wasHit/counters
                AngOps<R>::set_reset(s.ang[i], f, attacker);
                s.fl[i] = f;
            }
        }
    }
}

// One env.step() on registers.  Returns done, sets result; rew[] = per-agent reward.
template <int NG, int NA, typename R>
__device__ __forceinline__ bool step_env(Env<NG + NA, R> &s, const int (&act)[NG + NA], int max_steps,
                                         R (&rew)[NG + NA], int &result) {
    constexpr int A = NG + NA;
    typedef K<R> C;

    // ---- apply_laser_effect (core.py:254-302), on pre-move state and pre-kill alive flags -------
    uint32_t alive0 = 0;   // bit i = alive before this step
#pragma unroll
    for (int i = 0; i < A; ++i) {
        if (s.fl[i] & F_ALIVE) { alive0 |= 1u << i; s.fl[i] &= ~(F_HIT | F_WASHIT); }
    }
#pragma unroll
    for (int i = 0; i < A; ++i) {
        if ((act[i] == 7) && (alive0 >> i & 1u)) {
            R sn, cs;
            AngOps<R>::sincos_heading(s.ang[i], sn, cs);
            const R p1x = s.x[i] + C::SIZE * cs, p1y = s.y[i] + C::SIZE * sn;   // apex (core.py:376)
            const int j0 = (i < NG) ? NG : 0, j1 = (i < NG) ? A : NG;           // the other team
#pragma unroll
            for (int j = 0; j < A; ++j) {
                if (j >= j0 && j < j1 && (alive0 >> j & 1u)) {
                    // barycentric coordinates of q in the laser triangle, in the shooter's frame:
                    // u along the heading, v across; lambda1 = 1-a, lambda2,3 = (a +- b)/2
                    const R dx = s.x[j] - p1x, dy = s.y[j] - p1y;
                    const R a = (dx * cs + dy * sn) * C::INV_LC8;
                    const R b = (dy * cs - dx * sn) * C::INV_LS8;
                    if (a <= R(1) && a + b >= R(0) && a - b >= R(0)) {          // all lambda >= 0 (core.py:384-390)
                        uint32_t fi = s.fl[i] | F_HIT;
                        if (((fi >> F_NHIT_SHIFT) & F_CNT_MASK) < F_CNT_MASK) fi += 1u << F_NHIT_SHIFT;
                        s.fl[i] = fi;
                        uint32_t fj = s.fl[j] | F_WASHIT;
                        if (((fj >> F_NWAS_SHIFT) & F_CNT_MASK) < F_CNT_MASK) fj += 1u << F_NWAS_SHIFT;
                        s.fl[j] = fj;
                    }
                }
            }
        }
    }
    // dead agents lose justDied (core.py:287-289); alive agents that were hit die (core.py:293-302)
    uint32_t alive = 0;
#pragma unroll
    for (int i = 0; i < A; ++i) {
        uint32_t f = s.fl[i];
        if (!(f & F_ALIVE)) f &= ~F_JD;
        else if (f & F_WASHIT) f = (f & ~F_ALIVE) | F_JD;
        s.fl[i] = f;
        if (f & F_ALIVE) alive |= 1u << i;
    }

    // ---- forces on agents alive after the kill (core.py:204-210) --------------------------------
    R fx[A], fy[A];
#pragma unroll
    for (int i = 0; i < A; ++i) {   // _set_action + apply_action_force: u = accel * {0,+-1}
        const int a = act[i];
        fx[i] = a == 1 ? C::ACCEL : (a == 2 ? -C::ACCEL : R(0));
        fy[i] = a == 3 ? C::ACCEL : (a == 4 ? -C::ACCEL : R(0));
    }
#pragma unroll
    for (int a = 0; a < A; ++a) {   // apply_environment_force / get_collision_force (core.py:440-456)
#pragma unroll
        for (int b = a + 1; b < A; ++b) {
            const R dx = s.x[a] - s.x[b], dy = s.y[a] - s.y[b];
            const R r2 = dx * dx + dy * dy;
            if (r2 < ContactGate<R>::R2 && (alive >> a & 1u) && (alive >> b & 1u)) {
                const R rinv = rsqrt_t(r2);
                const R dist = r2 * rinv;
                const R g = C::CONTACT_FORCE * penetration(C::DIST_MIN - dist) * rinv;   // r2 == 0 -> NaN like the reference
                fx[a] += g * dx; fy[a] += g * dy;
                fx[b] -= g * dx; fy[b] -= g * dy;
            }
        }
    }

    // ---- wall force + integrate_state (core.py:246-252,459-472,305-338), alive agents only ------
#pragma unroll
    for (int i = 0; i < A; ++i) {
        if (alive >> i & 1u) {
            const R x = s.x[i], y = s.y[i];
            const R wx = C::CONTACT_FORCE * (penetration(-C::WALL_X - x) - penetration(x - C::WALL_X));
            const R wy = C::CONTACT_FORCE * (penetration(-C::WALL_Y - y) - penetration(y - C::WALL_Y));
            R vx = s.vx[i] * C::DAMP + (fx[i] + wx) * C::DT;
            R vy = s.vy[i] * C::DAMP + (fy[i] + wy) * C::DT;
            const R sp2 = vx * vx + vy * vy;
            if (sp2 > C::MAX_SPEED * C::MAX_SPEED) {
                const R k = C::MAX_SPEED * rsqrt_t(sp2);
                vx *= k; vy *= k;
            }
            AngOps<R>::advance(s.ang[i], s.fl[i], act[i]);
            s.vx[i] = vx; s.vy[i] = vy;
            s.x[i] = x + vx * C::DT; s.y[i] = y + vy * C::DT;
        }
    }

    // ---- reward (fortattack_env_v1.py:87-188) on post-move state ---------------------------------
    R d[A];
    int n_alive_att = 0;
    R min_att = R(1e30);
    bool reached = false;
#pragma unroll
    for (int i = 0; i < A; ++i) {
        const R ddx = s.x[i], ddy = s.y[i] - C::DOOR_Y;
        d[i] = sqrt_t(ddx * ddx + ddy * ddy);
        if (i >= NG && (alive >> i & 1u)) {
            n_alive_att += 1;
            min_att = min_t(min_att, d[i]);
        }
    }
    reached = min_att < C::FORT_DIM;
#pragma unroll
    for (int i = 0; i < A; ++i) {
        const uint32_t f = s.fl[i];
        R r = R(0);
        if (f & (F_ALIVE | F_JD)) {
            const R pd = s.pd[i];
            const bool has_prev = pd == pd;
            const bool shoot = act[i] == 7;
            if (i >= NG) {   // attacker_reward v1:94-128
                if (has_prev) r += R(2) * (pd - d[i]);
                if (d[i] < C::FORT_DIM) r += R(10);
                if (shoot) r -= R(1);
                if (f & F_HIT) r += R(3);
                if (f & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r -= R(10);
            } else {         // guard_reward v1:130-188
                if (has_prev) {
                    if (d[i] > C::GUARD_RING && pd <= C::GUARD_RING) r = R(-1);
                    else if (d[i] <= C::GUARD_RING && pd > C::GUARD_RING) r = R(1);
                }
                if (reached) r -= R(10);            // some alive attacker inside the fort
                if (shoot) r -= R(0.1);
                if (f & F_HIT) r += R(3);
                if (f & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r += R(10);
            }
            s.pd[i] = d[i];
        }
        rew[i] = r;
    }

    // ---- _get_done (fortattack.py:202-225) and time_step += 1 (:171) ----------------------------
    bool dn = true;
    if (reached) result = 3;
    else if (n_alive_att == 0) result = 1;
    else if (s.t == max_steps - 1) result = 2;
    else { result = 0; dn = false; }
    s.t += 1;
    return dn;
}

// ------------------------------------------------------------------------------------------------
// global <-> register movement.  Plane element (i, e) lives at index i*E + e: a warp's 32 lanes read
// 32 consecutive elements (512 B of pv, 256 B of ap, 128 B of fl/act) per instruction.
template <int A, typename R>
__device__ __forceinline__ void load_env(Env<A, R> &s, const StateView<R> &st, size_t E, int e) {
#pragma unroll
    for (int i = 0; i < A; ++i) {
        const typename VecT<R>::T4 v = st.pv[i * E + e];
        const typename VecT<R>::T2 w = st.ap[i * E + e];
        s.x[i] = v.x; s.y[i] = v.y; s.vx[i] = v.z; s.vy[i] = v.w;
        s.ang[i] = w.x; s.pd[i] = w.y;
        s.fl[i] = st.fl[i * E + e];
    }
    s.t = st.tstep[e];
}

template <int A, typename R>
__device__ __forceinline__ void store_env(const Env<A, R> &s, const StateView<R> &st, size_t E, int e) {
#pragma unroll
    for (int i = 0; i < A; ++i) {
        typename VecT<R>::T4 v; v.x = s.x[i]; v.y = s.y[i]; v.z = s.vx[i]; v.w = s.vy[i];
        typename VecT<R>::T2 w; w.x = s.ang[i]; w.y = s.pd[i];
        st.pv[i * E + e] = v;
        st.ap[i * E + e] = w;
        st.fl[i * E + e] = s.fl[i];
    }
    st.tstep[e] = s.t;
}

constexpr int MAX_WARPS = 4;          // threads per block <= 128
constexpr int OBS_DIM = 6;

// observation rows [alive, x, y, ang, vx, vy] (fortattack_env_v1.py:238) for plane set obs_t = [A][E][6].
// vec (warp-uniform): the warp is full and its 32 rows start 16-byte aligned -> transpose through the
// per-warp stage and write whole sectors with 16-byte stores; otherwise 3 row-local stores per lane.
template <int A, typename R>
__device__ __forceinline__ void store_obs(const Env<A, R> &s, R *obs_t, size_t E, int e, int lane, bool valid,
                                          bool vec, R (*stg)[32 * OBS_DIM]) {
    typedef typename VecT<R>::T2 T2;
    if (vec) __syncwarp();            // the stage may still be read by slower lanes (previous step)
#pragma unroll
    for (int i = 0; i < A; ++i) {
        T2 o0, o1, o2;
        o0.x = (s.fl[i] & F_ALIVE) ? R(1) : R(0); o0.y = s.x[i];
        o1.x = s.y[i]; o1.y = AngOps<R>::full(s.ang[i], s.fl[i]);
        o2.x = s.vx[i]; o2.y = s.vy[i];
        if (vec) {
            R *sb = stg[i & 1];
            T2 *w = reinterpret_cast<T2 *>(sb + lane * OBS_DIM);   // lane stride 24/48 B: conflict-free
            w[0] = o0; w[1] = o1; w[2] = o2;
            __syncwarp();
            const uint4 *src = reinterpret_cast<const uint4 *>(sb);
            uint4 *dst = reinterpret_cast<uint4 *>(obs_t + (i * E + (size_t)(e - lane)) * OBS_DIM);
            constexpr int NV = 32 * OBS_DIM * (int)sizeof(R) / 16;
#pragma unroll
            for (int q = lane; q < NV; q += 32) dst[q] = src[q];
        } else if (valid) {
            T2 *g = reinterpret_cast<T2 *>(obs_t + (i * E + (size_t)e) * OBS_DIM);
            g[0] = o0; g[1] = o1; g[2] = o2;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// The step kernel.  MANY = false: one env.step() per launch (fa_step).  MANY = true: T steps per
// launch with the state resident in registers and auto-reset always on (fa_step_many); the next
// step's actions are fetched before the current step is computed.
template <int NG, int NA, typename R, bool MANY>
__global__ void __launch_bounds__(32 * MAX_WARPS) fa_step_kernel(const StepParams<R> p) {
    constexpr int A = NG + NA;
    __shared__ __align__(16) R stage[MAX_WARPS][2][32 * OBS_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e - lane >= p.E) return;                       // whole warp out of range
    const bool valid = e < p.E;
    const int ec = valid ? e : p.E - 1;                // tail lanes shadow the last env, never store
    const size_t E = (size_t)p.E;
    const bool vec = p.obs_vec_ok && (e - lane + 32 <= p.E);

    pdl_launch_dependents();
    int act[A];
#pragma unroll
    for (int i = 0; i < A; ++i) act[i] = p.act[i * E + ec];
    pdl_wait();

    Env<A, R> s;
    load_env<A, R>(s, p.st, E, ec);
    uint32_t ep = 0;
    if (MANY) ep = p.st.episode[ec];

    const int T = MANY ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        int nxt[A];
        if (MANY && t + 1 < T) {
#pragma unroll
            for (int i = 0; i < A; ++i) nxt[i] = p.act[((size_t)(t + 1) * A + i) * E + ec];
        }
        R rew[A];
        int result;
        uint32_t alive_before = 0;
#pragma unroll
        for (int i = 0; i < A; ++i) alive_before |= (s.fl[i] & F_ALIVE) << i;
        const bool dn = step_env<NG, NA, R>(s, act, p.max_steps, rew, result);
        if (p.alive_end != nullptr && valid) {         // world.numAliveGuards / numAliveAttackers as the episode's last step leaves them
            int ag = 0, aa = 0;
#pragma unroll
            for (int i = 0; i < A; ++i) {
                if (s.fl[i] & F_ALIVE) { if (i < NG) ++ag; else ++aa; }
            }
            p.alive_end[(size_t)t * E + e] = (uint8_t)(ag | (aa << 4));
        }
        if (dn && (MANY || p.auto_reset)) {
            if (!MANY) ep = p.st.episode[ec];
            reset_env<NG, NA, R>(s, p.seed, p.env_id0 + (uint64_t)ec, ep);
            ep += 1;
            if (!MANY && valid) p.st.episode[e] = ep;
        }
        const size_t plane = (size_t)t * A;
        if (p.rew != nullptr && valid) {
#pragma unroll
            for (int i = 0; i < A; ++i) p.rew[(plane + i) * E + e] = rew[i];
        }
        if (p.done != nullptr && valid) p.done[(size_t)t * E + e] = dn ? 1 : 0;
        if (p.result != nullptr && valid) p.result[(size_t)t * E + e] = (uint8_t)result;
        if ((p.mask_next != nullptr || p.ep_rew != nullptr || p.end_next != nullptr) && valid) {
#pragma unroll
            for (int i = 0; i < A; ++i)
                rollout_outputs<R>(p, (size_t)t, A, E, i, e, alive_before >> i & 1u, s.fl[i] & F_ALIVE, dn, rew[i]);
        }
        if (p.obs != nullptr) store_obs<A, R>(s, p.obs + plane * E * OBS_DIM, E, e, lane, valid, vec, stage[warp]);
        if (MANY) {
#pragma unroll
            for (int i = 0; i < A; ++i) act[i] = nxt[i];
        }
    }
    if (valid) {
        store_env<A, R>(s, p.st, E, e);
        if (MANY) p.st.episode[e] = ep;
    }
}

// ------------------------------------------------------------------------------------------------
// The same step with ONE THREAD PER AGENT, for batches too small to fill the GPU with one thread per
// env (E=4096 is 128 warps on 592 warp schedulers; the per-env instruction chain, not bandwidth, is
// the limit there).  Block = A warps x 32 envs: warp i holds agent i of 32 consecutive envs, so global
// accesses stay exactly as coalesced as in the thread-per-env kernel; the per-env agent block needed by
// the O(A^2) laser and contact tests is staged in shared memory ([A][32] planes, conflict-free) across
// three block barriers per step:  positions/flags -> laser + kill -> forces/integrate -> reward/done.
template <int A, typename R> struct WideShared {
    R x0[A][32], y0[A][32];        // pre-move positions
    R cs[A][32], sn[A][32];        // heading of agents that shoot this step
    R d1[A][32];                   // post-move distance to the door
    uint32_t f0[A][32];            // bit 0 alive before the step, bit 1 shoots
    uint32_t a1[A][32];            // alive after the kill
};

template <int NG, int NA, typename R, bool MANY>
__global__ void __launch_bounds__(32 * (NG + NA)) fa_step_wide_kernel(const StepParams<R> p) {
    constexpr int A = NG + NA;
    constexpr int NOPP_MAX = NG > NA ? NG : NA;
    typedef K<R> C;
    __shared__ WideShared<A, R> sh;
    __shared__ __align__(16) R stage[A][2][32 * OBS_DIM];
    const int lane = threadIdx.x & 31, i = threadIdx.x >> 5;          // agent index = warp index
    const int e = blockIdx.x * 32 + lane;
    const bool valid = e < p.E;
    const int ec = valid ? e : p.E - 1;
    const size_t E = (size_t)p.E;
    const bool vec = p.obs_vec_ok && (blockIdx.x * 32 + 32 <= p.E);
    const bool attacker = i >= NG;                                     // warp-uniform
    const int j0 = attacker ? 0 : NG, nopp = attacker ? NG : NA;      // the other team

    pdl_launch_dependents();
    int act = p.act[i * E + ec];
    pdl_wait();
    Env<1, R> s;                                                       // this thread's agent
    {
        const typename VecT<R>::T4 v = p.st.pv[i * E + ec];
        const typename VecT<R>::T2 w = p.st.ap[i * E + ec];
        s.x[0] = v.x; s.y[0] = v.y; s.vx[0] = v.z; s.vy[0] = v.w; s.ang[0] = w.x; s.pd[0] = w.y;
        s.fl[0] = p.st.fl[i * E + ec];
        s.t = p.st.tstep[ec];
    }
    uint32_t ep = p.st.episode[ec];
    const int T = MANY ? p.T : 1;
    for (int t = 0; t < T; ++t) {
        int nxt = 0;
        if (MANY && t + 1 < T) nxt = p.act[((size_t)(t + 1) * A + i) * E + ec];
        R x = s.x[0], y = s.y[0];
        uint32_t fl = s.fl[0];
        // ---- stage the agent block --------------------------------------------------------------
        const bool alive0 = fl & F_ALIVE, shoot = act == 7;
        if (alive0) fl &= ~(F_HIT | F_WASHIT);
        R sn = R(0), cs = R(0);
        if (shoot && alive0) AngOps<R>::sincos_heading(s.ang[0], sn, cs);
        sh.x0[i][lane] = x; sh.y0[i][lane] = y; sh.cs[i][lane] = cs; sh.sn[i][lane] = sn;
        sh.f0[i][lane] = (alive0 ? 1u : 0u) | (shoot ? 2u : 0u);
        __syncthreads();
        // ---- apply_laser_effect: as shooter and as victim (core.py:254-302) ----------------------
        if (alive0) {
            const R p1x = x + C::SIZE * cs, p1y = y + C::SIZE * sn;
#pragma unroll
            for (int jj = 0; jj < NOPP_MAX; ++jj) {
                if (jj < nopp) {
                    const int j = j0 + jj;
                    const uint32_t fj = sh.f0[j][lane];
                    if (fj & 1u) {
                        const R xj = sh.x0[j][lane], yj = sh.y0[j][lane];
                        if (shoot) {
                            const R dx = xj - p1x, dy = yj - p1y;
                            const R a = (dx * cs + dy * sn) * C::INV_LC8, b = (dy * cs - dx * sn) * C::INV_LS8;
                            if (a <= R(1) && a + b >= R(0) && a - b >= R(0)) {
                                fl |= F_HIT;
                                if (((fl >> F_NHIT_SHIFT) & F_CNT_MASK) < F_CNT_MASK) fl += 1u << F_NHIT_SHIFT;
                            }
                        }
                        if (fj & 2u) {
                            const R cj = sh.cs[j][lane], sj = sh.sn[j][lane];
                            const R dx = x - (xj + C::SIZE * cj), dy = y - (yj + C::SIZE * sj);
                            const R a = (dx * cj + dy * sj) * C::INV_LC8, b = (dy * cj - dx * sj) * C::INV_LS8;
                            if (a <= R(1) && a + b >= R(0) && a - b >= R(0)) {
                                fl |= F_WASHIT;
                                if (((fl >> F_NWAS_SHIFT) & F_CNT_MASK) < F_CNT_MASK) fl += 1u << F_NWAS_SHIFT;
                            }
                        }
                    }
                }
            }
        }
        if (!alive0) fl &= ~F_JD;
        else if (fl & F_WASHIT) fl = (fl & ~F_ALIVE) | F_JD;
        const bool alive1 = fl & F_ALIVE;
        sh.a1[i][lane] = alive1 ? 1u : 0u;
        __syncthreads();
        // ---- forces on this agent + integrate (core.py:204-213) ---------------------------------
        if (alive1) {
            R fx = act == 1 ? C::ACCEL : (act == 2 ? -C::ACCEL : R(0));
            R fy = act == 3 ? C::ACCEL : (act == 4 ? -C::ACCEL : R(0));
#pragma unroll
            for (int j = 0; j < A; ++j) {
                if (j != i && sh.a1[j][lane]) {
                    const R dx = x - sh.x0[j][lane], dy = y - sh.y0[j][lane];
                    const R r2 = dx * dx + dy * dy;
                    if (r2 < ContactGate<R>::R2) {
                        const R rinv = rsqrt_t(r2);
                        const R g = C::CONTACT_FORCE * penetration(C::DIST_MIN - r2 * rinv) * rinv;
                        fx += g * dx; fy += g * dy;
                    }
                }
            }
            fx += C::CONTACT_FORCE * (penetration(-C::WALL_X - x) - penetration(x - C::WALL_X));
            fy += C::CONTACT_FORCE * (penetration(-C::WALL_Y - y) - penetration(y - C::WALL_Y));
            R vx = s.vx[0] * C::DAMP + fx * C::DT, vy = s.vy[0] * C::DAMP + fy * C::DT;
            const R sp2 = vx * vx + vy * vy;
            if (sp2 > C::MAX_SPEED * C::MAX_SPEED) {
                const R k = C::MAX_SPEED * rsqrt_t(sp2);
                vx *= k; vy *= k;
            }
            AngOps<R>::advance(s.ang[0], fl, act);
            s.vx[0] = vx; s.vy[0] = vy;
            x += vx * C::DT; y += vy * C::DT;
        }
        const R ddy = y - C::DOOR_Y;
        const R d = sqrt_t(x * x + ddy * ddy);
        sh.d1[i][lane] = d;
        __syncthreads();
        // ---- reward, done (fortattack_env_v1.py:87-188, fortattack.py:202-225) -------------------
        int n_alive_att = 0;
        R min_att = R(1e30);
#pragma unroll
        for (int j = NG; j < A; ++j) {
            if (sh.a1[j][lane]) { n_alive_att += 1; min_att = min_t(min_att, sh.d1[j][lane]); }
        }
        const bool reached = min_att < C::FORT_DIM;
        R r = R(0);
        if (fl & (F_ALIVE | F_JD)) {
            const R pd = s.pd[0];
            const bool has_prev = pd == pd;
            if (attacker) {
                if (has_prev) r += R(2) * (pd - d);
                if (d < C::FORT_DIM) r += R(10);
                if (shoot) r -= R(1);
                if (fl & F_HIT) r += R(3);
                if (fl & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r -= R(10);
            } else {
                if (has_prev) {
                    if (d > C::GUARD_RING && pd <= C::GUARD_RING) r = R(-1);
                    else if (d <= C::GUARD_RING && pd > C::GUARD_RING) r = R(1);
                }
                if (reached) r -= R(10);
                if (shoot) r -= R(0.1);
                if (fl & F_HIT) r += R(3);
                if (fl & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r += R(10);
            }
            s.pd[0] = d;
        }
        int result;
        bool dn = true;
        if (reached) result = 3;
        else if (n_alive_att == 0) result = 1;
        else if (s.t == p.max_steps - 1) result = 2;
        else { result = 0; dn = false; }
        s.t += 1;
        s.x[0] = x; s.y[0] = y; s.fl[0] = fl;
        if (dn && (MANY || p.auto_reset)) {
            // reset_world for this thread's agent: Philox block of its pair, lanes (2h, 2h+1)
            uint32_t c[4] = {(uint32_t)(p.env_id0 + (uint64_t)ec), (uint32_t)((p.env_id0 + (uint64_t)ec) >> 32), ep,
                             (uint32_t)(i >> 1)};
            philox4x32_10((uint32_t)p.seed, (uint32_t)(p.seed >> 32), c);
            const int h = i & 1;
            const double ux = u01(h ? c[2] : c[0]), uy = u01(h ? c[3] : c[1]);
            double px, py;
            if (attacker) {
                px = __dadd_rn(-1.0, __dmul_rn(1.0 - (-1.0), ux));
                py = __dadd_rn(-0.8, __dmul_rn(0.8 * -0.8 - (-0.8), uy));
            } else {
                const double lo = -0.8 * 0.15 / 2, hi = 0.8 * 0.15 / 2;
                px = __dadd_rn(lo, __dmul_rn(hi - lo, ux));
                py = __dadd_rn(0.8 * 0.8, __dmul_rn(0.8 - 0.8 * 0.8, uy));
            }
            s.x[0] = (R)px; s.y[0] = (R)py; s.vx[0] = R(0); s.vy[0] = R(0);
            uint32_t f = (s.fl[0] & (F_JD | (~0u << F_WRAP_SHIFT))) | F_ALIVE;
            AngOps<R>::set_reset(s.ang[0], f, attacker);
            s.fl[0] = f;
            s.t = 0;
            ep += 1;
        }
        // ---- outputs ----------------------------------------------------------------------------
        const size_t plane = (size_t)t * A;
        if (p.rew != nullptr && valid) p.rew[(plane + i) * E + e] = r;
        if (valid) rollout_outputs<R>(p, (size_t)t, A, E, i, e, alive0, s.fl[0] & F_ALIVE, dn, r);
        if (i == 0 && valid) {
            if (p.done != nullptr) p.done[(size_t)t * E + e] = dn ? 1 : 0;
            if (p.result != nullptr) p.result[(size_t)t * E + e] = (uint8_t)result;
            if (p.alive_end != nullptr) {
                int ag = 0, aa = 0;
#pragma unroll
                for (int j = 0; j < A; ++j) {
                    if (sh.a1[j][lane]) { if (j < NG) ++ag; else ++aa; }
                }
                p.alive_end[(size_t)t * E + e] = (uint8_t)(ag | (aa << 4));
            }
        }
        if (p.obs != nullptr) {
            typedef typename VecT<R>::T2 T2;
            T2 o0, o1, o2;
            o0.x = (s.fl[0] & F_ALIVE) ? R(1) : R(0); o0.y = s.x[0];
            o1.x = s.y[0]; o1.y = AngOps<R>::full(s.ang[0], s.fl[0]);
            o2.x = s.vx[0]; o2.y = s.vy[0];
            R *obs_t = p.obs + plane * E * OBS_DIM;
            if (vec) {
                R *sb = stage[i][t & 1];
                T2 *w = reinterpret_cast<T2 *>(sb + lane * OBS_DIM);
                w[0] = o0; w[1] = o1; w[2] = o2;
                __syncwarp();
                const uint4 *src = reinterpret_cast<const uint4 *>(sb);
                uint4 *dst = reinterpret_cast<uint4 *>(obs_t + (i * E + (size_t)(e - lane)) * OBS_DIM);
                constexpr int NV = 32 * OBS_DIM * (int)sizeof(R) / 16;
#pragma unroll
                for (int q = lane; q < NV; q += 32) dst[q] = src[q];
            } else if (valid) {
                T2 *g = reinterpret_cast<T2 *>(obs_t + (i * E + (size_t)e) * OBS_DIM);
                g[0] = o0; g[1] = o1; g[2] = o2;
            }
        }
        if (MANY) act = nxt;
    }
    if (valid) {
        typename VecT<R>::T4 v; v.x = s.x[0]; v.y = s.y[0]; v.z = s.vx[0]; v.w = s.vy[0];
        typename VecT<R>::T2 w; w.x = s.ang[0]; w.y = s.pd[0];
        p.st.pv[i * E + e] = v;
        p.st.ap[i * E + e] = w;
        p.st.fl[i * E + e] = s.fl[0];
        if (i == 0) { p.st.tstep[e] = s.t; p.st.episode[e] = ep; }
    }
}

// ------------------------------------------------------------------------------------------------
// The same step with one SUB-WARP GROUP per env: G = 2/4/8/16 consecutive lanes (the power of two >= A) hold the agents of
// one env, 32 / G envs per warp, and nothing crosses a warp -- no block barrier anywhere in the step.
//   agent block   each lane publishes (x, y, cos, sin) of its agent as ONE 16-byte shared-memory word in the warp's
//                 stage (double-buffered by step parity, one __syncwarp per step) and reads the A words of its env:
//                 the O(A^2) laser and contact tests then run on registers
//   masks         alive-before / shoots / alive-after-the-kill / inside-the-fort are warp ballots, shifted down to the
//                 group: "who was hit", "how many attackers are left" and "did the nearest alive attacker reach the
//                 fort" (min over the team < FORT_DIM  <=>  any member < FORT_DIM) are popc / != 0 on those masks
// Per-agent arithmetic is statement for statement that of fa_step_wide_kernel, so the two mappings agree bit for bit.
// Global accesses: a warp touches 32 / G consecutive envs of each agent plane (whole 32-byte sectors for pv / ap / obs); this
// mapping is chosen for batches that live in L2 (FA_MAP_AUTO: E < 37 888), where latency, not sector efficiency, is the limit.
template <int A> struct GroupOf { static constexpr int G = A <= 2 ? 2 : (A <= 4 ? 4 : (A <= 8 ? 8 : 16)); };

template <typename R>
__device__ __forceinline__ bool in_cone(R sx, R sy, R cs, R sn, R qx, R qy) {
    typedef K<R> C;
    const R dx = qx - (sx + C::SIZE * cs), dy = qy - (sy + C::SIZE * sn);     // q relative to the apex (core.py:376)
    const R a = (dx * cs + dy * sn) * C::INV_LC8, b = (dy * cs - dx * sn) * C::INV_LS8;
    return a <= R(1) && a + b >= R(0) && a - b >= R(0);
}

template <int NG, int NA, typename R, bool MANY>
__global__ void __launch_bounds__(32 * MAX_WARPS) fa_step_group_kernel(const StepParams<R> p) {
    constexpr int A = NG + NA, G = GroupOf<A>::G, EPW = 32 / G;
    constexpr uint32_t GM = (1u << A) - 1u, GRD_BITS = (1u << NG) - 1u, ATT_BITS = GM & ~GRD_BITS, FULL = 0xffffffffu;
    typedef K<R> C;
    typedef typename VecT<R>::T4 T4;
    typedef typename VecT<R>::T2 T2;
    __shared__ __align__(32) T4 block_stage[MAX_WARPS][2][32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int il = lane & (G - 1), base = lane - il;                   // lane within the group, first lane of the group
    const long long e0 = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * EPW;
    if (e0 >= p.E) return;                                             // whole warp out of range
    const int e = (int)e0 + lane / G;
    const bool agent = il < A;                                         // padding lanes (A < G) shadow the last agent, never store
    const bool valid = agent && e < p.E;
    const int ec = e < p.E ? e : p.E - 1, i = agent ? il : A - 1;
    const size_t E = (size_t)p.E;
    const bool attacker = i >= NG;
    const uint32_t opp_bits = attacker ? GRD_BITS : ATT_BITS;
    constexpr int NOPP_MAX = NG > NA ? NG : NA;
    const int j0 = attacker ? 0 : NG, nopp = attacker ? NG : NA;       // the other team

    pdl_launch_dependents();
    int act = p.act[i * E + ec], nxt1 = 0, nxt2 = 0;
    if (MANY && p.T > 1) nxt1 = p.act[((size_t)A + i) * E + ec];
    pdl_wait();
    R x, y, vx, vy, ang, pd;
    uint32_t fl;
    {
        const T4 v = p.st.pv[i * E + ec];
        const T2 w = p.st.ap[i * E + ec];
        x = v.x; y = v.y; vx = v.z; vy = v.w; ang = w.x; pd = w.y;
        fl = p.st.fl[i * E + ec];
    }
    int32_t ts = p.st.tstep[ec];
    uint32_t ep = p.st.episode[ec];
    const int T = MANY ? p.T : 1;
    size_t off_ae = (size_t)i * E + ec, off_e = (size_t)ec;            // this lane's element in [t][A][E] / [t][E] outputs
    for (int t = 0; t < T; ++t) {
        if (MANY && t + 2 < T) nxt2 = p.act[((size_t)(t + 2) * A + i) * E + ec];    // two steps ahead: a step is shorter than an L2 round trip
        // ---- publish the agent block ---------------------------------------------------------------
        // Everything from here to the reward is written as straight-line selects, not branches: with ~7 warps per SM at the
        // batch sizes this mapping serves, a step's duration is one warp's dependent-instruction chain, and the A cone tests
        // and A - 1 contact pairs of a lane are independent of each other -- the compiler can only overlap them if no
        // divergent region separates them.  (The selected values are those the branches of fa_step_wide_kernel compute.)
        const bool alive0 = agent && (fl & F_ALIVE), shoot = act == 7, shooter = shoot && alive0;
        fl = alive0 ? (fl & ~(F_HIT | F_WASHIT)) : fl;
        R sn, cs;
        AngOps<R>::sincos_heading(ang, sn, cs);
        sn = shooter ? sn : R(0); cs = shooter ? cs : R(0);
        T4 *stg = block_stage[warp][t & 1];
        {
            T4 me; me.x = x; me.y = y; me.z = cs; me.w = sn;
            stg[lane] = me;
        }
        const uint32_t m_alive0 = (__ballot_sync(FULL, alive0) >> base) & GM;
        const uint32_t m_shoot = (__ballot_sync(FULL, shooter) >> base) & GM;
        __syncwarp();
        // ---- apply_laser_effect: as shooter and as victim (core.py:254-302) ----------------------
        {
            const uint32_t targets = alive0 ? (opp_bits & m_alive0) : 0u;      // alive opponents of an alive agent
            const uint32_t as_shooter = (shoot ? targets : 0u) >> j0, as_victim = (targets & m_shoot) >> j0;
            int n_hit = 0, n_was = 0;
#pragma unroll
            for (int jj = 0; jj < NOPP_MAX; ++jj) {                            // opponent jj of this lane's team (lane-dependent address)
                const T4 q = stg[base + j0 + (jj < nopp ? jj : 0)];
                const bool h = in_cone<R>(x, y, cs, sn, q.x, q.y);
                const bool w = in_cone<R>(q.x, q.y, q.z, q.w, x, y);
                n_hit += (h && (as_shooter >> jj & 1u)) ? 1 : 0;
                n_was += (w && (as_victim >> jj & 1u)) ? 1 : 0;
            }
            const int c_hit = min((int)((fl >> F_NHIT_SHIFT) & F_CNT_MASK) + n_hit, (int)F_CNT_MASK);
            const int c_was = min((int)((fl >> F_NWAS_SHIFT) & F_CNT_MASK) + n_was, (int)F_CNT_MASK);
            fl = (fl & ~((F_CNT_MASK << F_NHIT_SHIFT) | (F_CNT_MASK << F_NWAS_SHIFT))) | ((uint32_t)c_hit << F_NHIT_SHIFT) |
                 ((uint32_t)c_was << F_NWAS_SHIFT) | (n_hit ? F_HIT : 0u) | (n_was ? F_WASHIT : 0u);
        }
        fl = !(fl & F_ALIVE) ? (fl & ~F_JD) : ((fl & F_WASHIT) ? ((fl & ~F_ALIVE) | F_JD) : fl);
        const bool alive1 = agent && (fl & F_ALIVE);
        const uint32_t m_alive1 = (__ballot_sync(FULL, alive1) >> base) & GM;
        // ---- forces on this agent + integrate (core.py:204-213) ---------------------------------
        {
            R fx = act == 1 ? C::ACCEL : (act == 2 ? -C::ACCEL : R(0));
            R fy = act == 3 ? C::ACCEL : (act == 4 ? -C::ACCEL : R(0));
            const uint32_t others = m_alive1 & ~(1u << i);
#pragma unroll
            for (int j = 0; j < A; ++j) {
                const T2 q = *reinterpret_cast<const T2 *>(&stg[base + j]);
                const R dx = x - q.x, dy = y - q.y;
                const R r2 = dx * dx + dy * dy;
                const bool touch = (others >> j & 1u) && r2 < ContactGate<R>::R2;
                if (sizeof(R) == 4) {
                    const R rinv = rsqrt_pinned(r2);
                    const R g = touch ? C::CONTACT_FORCE * penetration(C::DIST_MIN - r2 * rinv) * rinv : R(0);
                    fx += g * dx; fy += g * dy;
                } else if (touch) {                                   // double: the softplus itself (log1p / exp), rarely needed
                    const R rinv = rsqrt_t(r2);
                    const R g = C::CONTACT_FORCE * penetration(C::DIST_MIN - r2 * rinv) * rinv;
                    fx += g * dx; fy += g * dy;
                }
            }
            fx += C::CONTACT_FORCE * (penetration(-C::WALL_X - x) - penetration(x - C::WALL_X));
            fy += C::CONTACT_FORCE * (penetration(-C::WALL_Y - y) - penetration(y - C::WALL_Y));
            R nvx = vx * C::DAMP + fx * C::DT, nvy = vy * C::DAMP + fy * C::DT;
            const R sp2 = nvx * nvx + nvy * nvy;
            const R k = sp2 > C::MAX_SPEED * C::MAX_SPEED ? C::MAX_SPEED * rsqrt_t(sp2) : R(1);
            nvx = sp2 > C::MAX_SPEED * C::MAX_SPEED ? nvx * k : nvx;
            nvy = sp2 > C::MAX_SPEED * C::MAX_SPEED ? nvy * k : nvy;
            if (alive1) {
                AngOps<R>::advance(ang, fl, act);
                vx = nvx; vy = nvy;
                x += nvx * C::DT; y += nvy * C::DT;
            }
        }
        const R ddy = y - C::DOOR_Y;
        const R d = sqrt_t(x * x + ddy * ddy);
        // ---- reward, done (fortattack_env_v1.py:87-188, fortattack.py:202-225) -------------------
        const bool reached = ((__ballot_sync(FULL, alive1 && attacker && d < C::FORT_DIM) >> base) & GM) != 0u;
        const int n_alive_att = __popc(m_alive1 & ATT_BITS);
        R r = R(0);
        if (fl & (F_ALIVE | F_JD)) {
            const bool has_prev = pd == pd;
            if (attacker) {
                if (has_prev) r += R(2) * (pd - d);
                if (d < C::FORT_DIM) r += R(10);
                if (shoot) r -= R(1);
                if (fl & F_HIT) r += R(3);
                if (fl & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r -= R(10);
            } else {
                if (has_prev) {
                    if (d > C::GUARD_RING && pd <= C::GUARD_RING) r = R(-1);
                    else if (d <= C::GUARD_RING && pd > C::GUARD_RING) r = R(1);
                }
                if (reached) r -= R(10);
                if (shoot) r -= R(0.1);
                if (fl & F_HIT) r += R(3);
                if (fl & F_WASHIT) r -= R(3);
                if (n_alive_att == 0) r += R(10);
            }
            pd = d;
        }
        int result;
        bool dn = true;
        if (reached) result = 3;
        else if (n_alive_att == 0) result = 1;
        else if (ts == p.max_steps - 1) result = 2;
        else { result = 0; dn = false; }
        ts += 1;
        if (dn && (MANY || p.auto_reset)) {
            // reset_world for this lane's agent: Philox block of its pair, words (2h, 2h+1)
            const uint64_t id = p.env_id0 + (uint64_t)ec;
            uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), ep, (uint32_t)(i >> 1)};
            philox4x32_10((uint32_t)p.seed, (uint32_t)(p.seed >> 32), c);
            const int h = i & 1;
            const double ux = u01(h ? c[2] : c[0]), uy = u01(h ? c[3] : c[1]);
            double px, py;
            if (attacker) {
                px = __dadd_rn(-1.0, __dmul_rn(1.0 - (-1.0), ux));
                py = __dadd_rn(-0.8, __dmul_rn(0.8 * -0.8 - (-0.8), uy));
            } else {
                const double lo = -0.8 * 0.15 / 2, hi = 0.8 * 0.15 / 2;
                px = __dadd_rn(lo, __dmul_rn(hi - lo, ux));
                py = __dadd_rn(0.8 * 0.8, __dmul_rn(0.8 - 0.8 * 0.8, uy));
            }
            x = (R)px; y = (R)py; vx = R(0); vy = R(0);
            fl = (fl & (F_JD | (~0u << F_WRAP_SHIFT))) | F_ALIVE;
            AngOps<R>::set_reset(ang, fl, attacker);
            ts = 0;
            ep += 1;
        }
        // ---- outputs (element offsets of this step advance by one plane set per step) -------------
        if (valid) {
            if (p.rew != nullptr) p.rew[off_ae] = r;
            if (p.mask_next != nullptr) p.mask_next[off_ae] = ((dn ? (fl & F_ALIVE) != 0u : alive0)) ? 1.0f : 0.0f;
            if (p.ep_rew != nullptr && alive0) p.ep_rew[(size_t)i * E + e] += (float)r;
            if (i == 0) {
                if (p.end_next != nullptr) p.end_next[off_e] = dn ? 1 : 0;
                if (p.done != nullptr) p.done[off_e] = dn ? 1 : 0;
                if (p.result != nullptr) p.result[off_e] = (uint8_t)result;
                if (p.alive_end != nullptr)
                    p.alive_end[off_e] = (uint8_t)(__popc(m_alive1 & GRD_BITS) | (__popc(m_alive1 & ATT_BITS) << 4));
            }
            if (p.obs != nullptr) {
                T2 o0, o1, o2;
                o0.x = (fl & F_ALIVE) ? R(1) : R(0); o0.y = x;
                o1.x = y; o1.y = AngOps<R>::full(ang, fl);
                o2.x = vx; o2.y = vy;
                T2 *gdst = reinterpret_cast<T2 *>(p.obs + off_ae * OBS_DIM);
                gdst[0] = o0; gdst[1] = o1; gdst[2] = o2;
            }
        }
        off_ae += (size_t)A * E; off_e += E;
        if (MANY) { act = nxt1; nxt1 = nxt2; }
    }
    if (valid) {
        T4 v; v.x = x; v.y = y; v.z = vx; v.w = vy;
        T2 w; w.x = ang; w.y = pd;
        p.st.pv[i * E + e] = v;
        p.st.ap[i * E + e] = w;
        p.st.fl[i * E + e] = fl;
        if (i == 0) { p.st.tstep[e] = ts; p.st.episode[e] = ep; }
    }
}

// fa_reset: reset the masked envs (all if mask == nullptr) and report every env's observation.
template <int NG, int NA, typename R>
__global__ void __launch_bounds__(32 * MAX_WARPS)
fa_reset_kernel(StateView<R> st, const uint8_t *mask, R *obs, int nE, uint64_t seed, uint64_t env_id0) {
    constexpr int A = NG + NA;
    __shared__ __align__(16) R stage[MAX_WARPS][2][32 * OBS_DIM];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e - lane >= nE) return;
    const bool valid = e < nE;
    const int ec = valid ? e : nE - 1;
    const size_t E = (size_t)nE;
    Env<A, R> s;
    load_env<A, R>(s, st, E, ec);
    if (mask == nullptr || mask[ec]) {
        const uint32_t ep = st.episode[ec];
        reset_env<NG, NA, R>(s, seed, env_id0 + (uint64_t)ec, ep);
        if (valid) {
            store_env<A, R>(s, st, E, e);
            st.episode[e] = ep + 1;
        }
    }
    if (obs != nullptr) store_obs<A, R>(s, obs, E, e, lane, valid, false, stage[warp]);
}

// ------------------------------------------------------------------------------------------------
// layout-independent helpers (runtime A; not on the hot path)
template <typename R>
__global__ void fa_init_kernel(StateView<R> st, int nE, int A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    for (int i = 0; i < A; ++i) {
        const size_t k = (size_t)i * nE + e;
        typename VecT<R>::T4 v; v.x = v.y = v.z = v.w = R(0);
        typename VecT<R>::T2 w; w.x = R(0); w.y = (R)CUDART_NAN;     // prevDist = None (core.py:104)
        st.pv[k] = v; st.ap[k] = w; st.fl[k] = F_ALIVE;
    }
    st.tstep[e] = 0;
    st.episode[e] = 0;
}

// canonical state (include/fortattack.h FaState, env-major float64) -> internal planes
template <typename R>
__global__ void fa_set_state_kernel(StateView<R> st, const double *st_f, const uint8_t *st_i, const int32_t *tstep,
                                    const uint32_t *episode, int nE, int A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    for (int i = 0; i < A; ++i) {
        const double *f = st_f + ((size_t)e * A + i) * 6;
        const uint8_t *b = st_i + ((size_t)e * A + i) * 6;
        const size_t k = (size_t)i * nE + e;
        typename VecT<R>::T4 v; v.x = (R)f[0]; v.y = (R)f[1]; v.z = (R)f[2]; v.w = (R)f[3];
        uint32_t fl = (b[0] ? F_ALIVE : 0u) | (b[1] ? F_JD : 0u) | (b[2] ? F_HIT : 0u) | (b[3] ? F_WASHIT : 0u);
        fl |= (uint32_t)(b[4] < 15 ? b[4] : 15) << F_NHIT_SHIFT;
        fl |= (uint32_t)(b[5] < 15 ? b[5] : 15) << F_NWAS_SHIFT;
        typename VecT<R>::T2 w;
        if (sizeof(R) == 4) {
            double turns = floor(f[4] / 6.283185307179586);
            turns = turns < 0.0 ? 0.0 : (turns > 65535.0 ? 65535.0 : turns);
            w.x = (R)(f[4] - turns * 6.283185307179586);
            fl |= (uint32_t)turns << F_WRAP_SHIFT;
        } else {
            w.x = (R)f[4];
        }
        w.y = (R)f[5];
        st.pv[k] = v; st.ap[k] = w; st.fl[k] = fl;
    }
    st.tstep[e] = tstep[e];
    st.episode[e] = episode[e];
}

template <typename R>
__global__ void fa_get_state_kernel(StateView<R> st, double *st_f, uint8_t *st_i, int32_t *tstep, uint32_t *episode,
                                    int nE, int A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    for (int i = 0; i < A; ++i) {
        double *f = st_f + ((size_t)e * A + i) * 6;
        uint8_t *b = st_i + ((size_t)e * A + i) * 6;
        const size_t k = (size_t)i * nE + e;
        const typename VecT<R>::T4 v = st.pv[k];
        const typename VecT<R>::T2 w = st.ap[k];
        const uint32_t fl = st.fl[k];
        f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
        f[4] = sizeof(R) == 4 ? (double)w.x + (double)(fl >> F_WRAP_SHIFT) * 6.283185307179586 : (double)w.x;
        f[5] = w.y;
        b[0] = fl & F_ALIVE ? 1 : 0; b[1] = fl & F_JD ? 1 : 0; b[2] = fl & F_HIT ? 1 : 0; b[3] = fl & F_WASHIT ? 1 : 0;
        b[4] = (fl >> F_NHIT_SHIFT) & F_CNT_MASK; b[5] = (fl >> F_NWAS_SHIFT) & F_CNT_MASK;
    }
    tstep[e] = st.tstep[e];
    episode[e] = st.episode[e];
}

// world.numAliveGuards / numAliveAttackers: counts [2][E]
__global__ void fa_alive_counts_kernel(const uint32_t *fl, int32_t *counts, int nE, int n_guards, int A);

}  // namespace fa
