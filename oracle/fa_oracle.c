/*
 * fa_oracle.c -- CPU ORACLE for the FortAttack step path.  TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
 * load this library, and only as the checker or the timed CPU baseline.  The product
 * (emergent-multiagent-strategies_b200/) never links, imports or calls it.
 *
 * It is a plain-C, scalar, float64 restatement of the reference's algorithm, one environment at a
 * time, in the reference's own evaluation order (the reference is Python/numpy float64):
 *
 *   decode        gym_fortattack/fortattack.py:235-302   (_set_action)
 *   laser         gym_fortattack/core.py:254-302,373-390 (apply_laser_effect, get_tri_pts_arr, laser_hit)
 *   action force  gym_fortattack/core.py:221-228
 *   contact       gym_fortattack/core.py:231-243,440-456 (get_collision_force, softplus k=1e-10)
 *   wall          gym_fortattack/core.py:246-252,459-472
 *   integrate     gym_fortattack/core.py:305-338
 *   observation   gym_fortattack/envs/fortattack_env_v1.py:191-238
 *   reward        gym_fortattack/envs/fortattack_env_v1.py:87-188
 *   done          gym_fortattack/fortattack.py:202-225, :171 (time_step += 1)
 *   reset         gym_fortattack/envs/fortattack_env_v1.py:47-75
 *
 * Parity pin: tests/test_oracle_golden.py replays tests/golden/env_*.npz -- transitions produced by
 * running the unchanged reference (tests/golden/make_env_golden.py) -- through fa_oracle_step and
 * requires |d obs|,|d reward| <= 1e-12 and identical alive/justDied/hit/wasHit/done/result; and the
 * reference's own recorded trajectory out_files/1.npy (tests/golden/ref_traj_5v5.npy).
 *
 * Two deliberate, documented deviations from the literal reference text:
 *  (1) laser_hit solves the 3x3 barycentric system in closed form instead of numpy's SVD
 *      pseudo-inverse (core.py:365-371).  The laser triangle is never degenerate (fixed shape,
 *      area 0.226), so both give A^-1 b to ~1e-15; `margin` reports min|lambda| so tests can
 *      exclude boundary cases.
 *  (2) reset draws positions from Philox4x32-10 keyed by (seed, global env id, episode) instead of
 *      numpy's global MT19937 (v1:66,70): the stream cannot be shared with a GPU.  Same
 *      distribution, same draw order (x then y, agents in index order).
 *
 * State layout (mirrors tests/golden/env_*.npz):
 *   st_f [E][A][6] double : x, y, vx, vy, ang, prevDist (NaN = None)
 *   st_i [E][A][6] uint8  : alive, justDied, hit, wasHit, numHit, numWasHit
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#define FA_MAX_AGENTS 64

/* constants: core.py:32,100-101,115-128 ; fortattack_env_v1.py:16-17,33-35 */
static const double SIZE = 0.05, SHOOT_RAD = 0.8, DT = 0.1, DAMPING = 0.25;
static const double CONTACT_FORCE = 1e+2, CONTACT_MARGIN = 1e-10;
static const double WALL_XMIN = -1, WALL_XMAX = 1, WALL_YMIN = -0.8, WALL_YMAX = 0.8;
static const double FORT_DIM = 0.15, DOOR_X = 0, DOOR_Y = 0.8;
static const double ACCEL = 3, MAX_SPEED = 3, MAX_ROT = 0.17;

/* np.logaddexp(0, t) */
static double logaddexp0(double t) {
    if (t == 0.0) return log(2.0);
    double tmp = 0.0 - t;
    if (tmp > 0) return 0.0 + log1p(exp(-tmp));
    return t + log1p(exp(tmp));
}

/* Python float modulo u % w for w > 0 (core.py:336) */
static double pymod(double u, double w) {
    double m = fmod(u, w);
    if (m != 0.0 && m < 0) m += w;
    return m;
}

/* ---------------------------------------------------------------- Philox4x32-10 ------------ */
static void philox4x32_10(uint32_t k0, uint32_t k1, uint32_t c[4]) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}
static double u01(uint32_t r) { return ((double)r + 0.5) * (1.0 / 4294967296.0); }

/* reset_world: fortattack_env_v1.py:47-75.  prevDist / justDied are NOT reset. */
static void reset_env(int n_guards, int A, uint64_t seed, uint64_t env_id, uint32_t episode,
                      double *f, uint8_t *s, int32_t *tstep) {
    *tstep = 0;
    for (int pair = 0; pair * 2 < A; ++pair) {
        uint32_t c[4] = {(uint32_t)env_id, (uint32_t)(env_id >> 32), episode, (uint32_t)pair};
        philox4x32_10((uint32_t)seed, (uint32_t)(seed >> 32), c);
        for (int h = 0; h < 2; ++h) {
            int i = pair * 2 + h;
            if (i >= A) break;
            double ux = u01(c[2 * h]), uy = u01(c[2 * h + 1]);
            double *a = f + 6 * i;
            uint8_t *b = s + 6 * i;
            int attacker = i >= n_guards;
            b[0] = 1;                       /* alive */
            a[2] = 0.0; a[3] = 0.0;         /* p_vel */
            a[4] = attacker ? M_PI / 2 : 3 * M_PI / 2;
            if (attacker) {
                a[0] = WALL_XMIN + (WALL_XMAX - WALL_XMIN) * ux;
                a[1] = WALL_YMIN + (0.8 * WALL_YMIN - WALL_YMIN) * uy;
            } else {
                double lo = -0.8 * FORT_DIM / 2, hi = 0.8 * FORT_DIM / 2;
                a[0] = lo + (hi - lo) * ux;
                a[1] = 0.8 * WALL_YMAX + (WALL_YMAX - 0.8 * WALL_YMAX) * uy;
            }
            b[2] = 0; b[3] = 0; b[4] = 0; b[5] = 0;   /* hit, wasHit, numHit, numWasHit */
        }
    }
}

static void write_obs(int A, const double *f, const uint8_t *s, double *obs) {
    /* fortattack_env_v1.py:238 : [alive, x, y, ang, vx, vy] */
    for (int i = 0; i < A; ++i) {
        obs[6 * i + 0] = s[6 * i + 0] ? 1.0 : 0.0;
        obs[6 * i + 1] = f[6 * i + 0];
        obs[6 * i + 2] = f[6 * i + 1];
        obs[6 * i + 3] = f[6 * i + 4];
        obs[6 * i + 4] = f[6 * i + 2];
        obs[6 * i + 5] = f[6 * i + 3];
    }
}

/* ---------------------------------------------------------------- one env step -------------- */
static void step_env(int n_guards, int A, int max_steps, double *f, uint8_t *s, int32_t *tstep,
                     const int32_t *act, double *obs, double *rew, uint8_t *done, uint8_t *result,
                     double *margin_out) {
    double ux[FA_MAX_AGENTS], uy[FA_MAX_AGENTS], rot[FA_MAX_AGENTS];
    double fx[FA_MAX_AGENTS], fy[FA_MAX_AGENTS];
    uint8_t shoot[FA_MAX_AGENTS];
    double margin = INFINITY;

#define X(i) f[6 * (i) + 0]
#define Y(i) f[6 * (i) + 1]
#define VX(i) f[6 * (i) + 2]
#define VY(i) f[6 * (i) + 3]
#define ANG(i) f[6 * (i) + 4]
#define PD(i) f[6 * (i) + 5]
#define ALIVE(i) s[6 * (i) + 0]
#define JD(i) s[6 * (i) + 1]
#define HIT(i) s[6 * (i) + 2]
#define WASHIT(i) s[6 * (i) + 3]
#define NHIT(i) s[6 * (i) + 4]
#define NWASHIT(i) s[6 * (i) + 5]
#define ATTACKER(i) ((i) >= n_guards)

    /* _set_action, every agent incl. dead: fortattack.py:253-263,283-286 */
    for (int i = 0; i < A; ++i) {
        int a = act[i];
        ux[i] = 0; uy[i] = 0; rot[i] = 0;
        if (a == 1) ux[i] = +1.0;
        if (a == 2) ux[i] = -1.0;
        if (a == 3) uy[i] = +1.0;
        if (a == 4) uy[i] = -1.0;
        if (a == 5) rot[i] = +MAX_ROT;
        if (a == 6) rot[i] = -MAX_ROT;
        shoot[i] = (a == 7);
        ux[i] *= ACCEL; uy[i] *= ACCEL;
    }

    /* apply_laser_effect: core.py:254-302 */
    for (int i = 0; i < A; ++i) if (ALIVE(i)) { HIT(i) = 0; WASHIT(i) = 0; }
    for (int i = 0; i < A; ++i) {
        if (!ALIVE(i) || !shoot[i]) continue;
        /* get_tri_pts_arr: core.py:373-382 (shootWin = pi/4) */
        double ang = ANG(i), win = M_PI / 4;
        double p1x = X(i) + SIZE * cos(ang), p1y = Y(i) + SIZE * sin(ang);
        double p2x = p1x + SHOOT_RAD * cos(ang + win / 2), p2y = p1y + SHOOT_RAD * sin(ang + win / 2);
        double p3x = p1x + SHOOT_RAD * cos(ang - win / 2), p3y = p1y + SHOOT_RAD * sin(ang - win / 2);
        double det = (p2y - p3y) * (p1x - p3x) + (p3x - p2x) * (p1y - p3y);
        for (int b = 0; b < A; ++b) {
            if (!ALIVE(b) || ATTACKER(b) == ATTACKER(i)) continue;
            /* laser_hit: lambda = [p1 p2 p3; 1 1 1]^-1 [q; 1], hit iff all lambda >= 0 */
            double qx = X(b), qy = Y(b);
            double l1 = ((p2y - p3y) * (qx - p3x) + (p3x - p2x) * (qy - p3y)) / det;
            double l2 = ((p3y - p1y) * (qx - p3x) + (p1x - p3x) * (qy - p3y)) / det;
            double l3 = 1.0 - l1 - l2;
            double m = fmin(fabs(l1), fmin(fabs(l2), fabs(l3)));
            if (m < margin) margin = m;
            if (l1 >= 0 && l2 >= 0 && l3 >= 0) {
                HIT(i) = 1; if (NHIT(i) < 255) NHIT(i)++;
                WASHIT(b) = 1; if (NWASHIT(b) < 255) NWASHIT(b)++;
            }
        }
    }
    for (int i = 0; i < A; ++i) if (!ALIVE(i)) JD(i) = 0;            /* core.py:287-289 */
    for (int i = 0; i < A; ++i) if (ALIVE(i) && WASHIT(i)) {         /* core.py:293-302 */
        ALIVE(i) = 0; JD(i) = 1;
    }

    /* forces on agents alive after the kill: core.py:204-210 */
    for (int i = 0; i < A; ++i) { fx[i] = ux[i]; fy[i] = uy[i]; }    /* apply_action_force */
    for (int a = 0; a < A; ++a) {                                    /* apply_environment_force */
        if (!ALIVE(a)) continue;
        for (int b = a + 1; b < A; ++b) {
            if (!ALIVE(b)) continue;
            /* get_collision_force: core.py:440-456 */
            double dx = X(a) - X(b), dy = Y(a) - Y(b);
            double dist = sqrt(dx * dx + dy * dy);
            double dist_min = SIZE + SIZE;
            double k = CONTACT_MARGIN;
            double pen = logaddexp0(-(dist - dist_min) / k) * k;
            double gx = CONTACT_FORCE * dx / dist * pen, gy = CONTACT_FORCE * dy / dist * pen;
            fx[a] = gx + fx[a]; fy[a] = gy + fy[a];
            fx[b] = -gx + fx[b]; fy[b] = -gy + fy[b];
        }
    }
    for (int a = 0; a < A; ++a) {                                    /* apply_wall_collision_force */
        if (!ALIVE(a)) continue;
        double k = CONTACT_MARGIN;
        double d0 = X(a) - SIZE - WALL_XMIN, d1 = WALL_XMAX - X(a) - SIZE;
        double d2 = Y(a) - SIZE - WALL_YMIN, d3 = WALL_YMAX - Y(a) - SIZE;
        double f0 = CONTACT_FORCE * (logaddexp0(-d0 / k) * k), f1 = CONTACT_FORCE * (logaddexp0(-d1 / k) * k);
        double f2 = CONTACT_FORCE * (logaddexp0(-d2 / k) * k), f3 = CONTACT_FORCE * (logaddexp0(-d3 / k) * k);
        fx[a] = (f0 - f1) + fx[a];
        fy[a] = (f2 - f3) + fy[a];
    }

    /* integrate_state: core.py:324-338 (mass 1) */
    for (int i = 0; i < A; ++i) {
        if (!ALIVE(i)) continue;
        VX(i) = VX(i) * (1 - DAMPING); VY(i) = VY(i) * (1 - DAMPING);
        VX(i) += (fx[i] / 1.0) * DT;   VY(i) += (fy[i] / 1.0) * DT;
        double speed = sqrt(VX(i) * VX(i) + VY(i) * VY(i));
        if (speed > MAX_SPEED) {
            double n = sqrt(VX(i) * VX(i) + VY(i) * VY(i));
            VX(i) = VX(i) / n * MAX_SPEED; VY(i) = VY(i) / n * MAX_SPEED;
        }
        ANG(i) += pymod(rot[i], 2 * M_PI);
        X(i) += VX(i) * DT; Y(i) += VY(i) * DT;
    }

    /* observation + reward per agent: fortattack.py:158-163 */
    write_obs(A, f, s, obs);
    int n_alive_att = 0;
    for (int i = n_guards; i < A; ++i) n_alive_att += ALIVE(i);
    for (int i = 0; i < A; ++i) {
        double r = 0;
        if (ALIVE(i) || JD(i)) {                                     /* v1:87-92 */
            double ddx = X(i) - DOOR_X, ddy = Y(i) - DOOR_Y;
            double d = sqrt(ddx * ddx + ddy * ddy);
            int has_prev = !isnan(PD(i));
            if (ATTACKER(i)) {                                       /* attacker_reward v1:94-128 */
                double r0 = 0, r1 = 0, r2 = 0, r3 = 0, r4 = 0, r5 = 0;
                if (has_prev) r0 = 2 * (PD(i) - d);
                if (d < FORT_DIM) r1 = 10;
                if (shoot[i]) r2 = -1;
                if (HIT(i)) r3 = +3;
                if (WASHIT(i)) r4 = -3;
                if (n_alive_att == 0) r5 = -10;
                r = r0 + r1 + r2 + r3 + r4 + r5;
            } else {                                                 /* guard_reward v1:130-188 */
                double r0 = 0, r3 = 0, r4 = 0, r5 = 0, r6 = 0, r7 = 0;
                if (has_prev) {
                    if (d > 0.3 && PD(i) <= 0.3) r0 = -1;
                    else if (d <= 0.3 && PD(i) > 0.3) r0 = 1;
                }
                if (n_alive_att != 0) {
                    double mind = INFINITY;
                    for (int j = n_guards; j < A; ++j) if (ALIVE(j)) {
                        double ex = X(j) - DOOR_X, ey = Y(j) - DOOR_Y;
                        double dj = sqrt(ex * ex + ey * ey);
                        if (dj < mind) mind = dj;
                    }
                    if (mind < FORT_DIM) r3 = -10;
                }
                if (shoot[i]) r4 = -0.1;
                if (HIT(i)) r5 = 3;
                if (WASHIT(i)) r6 = -3;
                if (n_alive_att == 0) r7 = 10;
                r = r0 + 0 + 0 + r3 + r4 + r5 + r6 + r7 + 0;         /* rew0+rew1+...+rew8 */
            }
            PD(i) = d;
        }
        rew[i] = r;
    }

    /* _get_done: fortattack.py:202-225 */
    int dn = 0, res = 0;
    for (int j = n_guards; j < A && !dn; ++j) if (ALIVE(j)) {
        double ex = X(j) - DOOR_X, ey = Y(j) - DOOR_Y;
        if (sqrt(ex * ex + ey * ey) < FORT_DIM) { dn = 1; res = 3; }
    }
    if (!dn) {
        if (n_alive_att == 0) { dn = 1; res = 1; }
        else if (*tstep == max_steps - 1) { dn = 1; res = 2; }
    }
    *tstep += 1;                                                     /* fortattack.py:171 */
    *done = (uint8_t)dn; *result = (uint8_t)res;
    if (margin_out) *margin_out = margin;
}

/* ---------------------------------------------------------------- exported ------------------ */
int fa_oracle_max_agents(void) { return FA_MAX_AGENTS; }

/* Reset the envs with mask[e] != 0 (all if mask == NULL) and write their observations. */
int fa_oracle_reset(int E, int n_guards, int n_attackers, uint64_t seed, uint64_t env_id0,
                    const uint8_t *mask, double *st_f, uint8_t *st_i, int32_t *time_step,
                    uint32_t *episode, double *obs) {
    int A = n_guards + n_attackers;
    if (A > FA_MAX_AGENTS || A < 1) return -1;
    for (int e = 0; e < E; ++e) {
        if (mask && !mask[e]) continue;
        reset_env(n_guards, A, seed, env_id0 + (uint64_t)e, episode[e], st_f + (size_t)e * A * 6,
                  st_i + (size_t)e * A * 6, time_step + e);
        episode[e] += 1;
        if (obs) write_obs(A, st_f + (size_t)e * A * 6, st_i + (size_t)e * A * 6, obs + (size_t)e * A * 6);
    }
    return 0;
}

/* ---- host threading: contiguous env ranges, one pthread per range (no OpenMP in this image) -- */
typedef struct {
    int T, E, e0, e1, n_guards, A, max_steps, auto_reset;
    double *st_f; uint8_t *st_i; int32_t *time_step; const int32_t *actions;
    double *obs, *rew; uint8_t *done, *result; double *margin;
    uint64_t seed, env_id0; uint32_t *episode;
} Job;

static void run_step(const Job *j) {
    int A = j->A;
    for (int e = j->e0; e < j->e1; ++e) {
        size_t o = (size_t)e * A;
        step_env(j->n_guards, A, j->max_steps, j->st_f + o * 6, j->st_i + o * 6, j->time_step + e,
                 j->actions + o, j->obs + o * 6, j->rew + o, j->done + e, j->result + e,
                 j->margin ? j->margin + e : 0);
        if (j->auto_reset && j->done[e]) {
            reset_env(j->n_guards, A, j->seed, j->env_id0 + (uint64_t)e, j->episode[e], j->st_f + o * 6,
                      j->st_i + o * 6, j->time_step + e);
            j->episode[e] += 1;
            write_obs(A, j->st_f + o * 6, j->st_i + o * 6, j->obs + o * 6);
        }
    }
}

static void run_step_many(const Job *j) {
    int A = j->A, E = j->E;
    for (int e = j->e0; e < j->e1; ++e) {
        size_t o = (size_t)e * A;
        double lobs[FA_MAX_AGENTS * 6], lrew[FA_MAX_AGENTS];
        uint8_t ld, lr;
        for (int t = 0; t < j->T; ++t) {
            size_t to = ((size_t)t * E + e) * A;
            step_env(j->n_guards, A, j->max_steps, j->st_f + o * 6, j->st_i + o * 6, j->time_step + e,
                     j->actions + to, lobs, lrew, &ld, &lr, 0);
            if (ld) {
                reset_env(j->n_guards, A, j->seed, j->env_id0 + (uint64_t)e, j->episode[e],
                          j->st_f + o * 6, j->st_i + o * 6, j->time_step + e);
                j->episode[e] += 1;
                write_obs(A, j->st_f + o * 6, j->st_i + o * 6, lobs);
            }
            if (j->obs) memcpy(j->obs + to * 6, lobs, sizeof(double) * A * 6);
            if (j->rew) memcpy(j->rew + to, lrew, sizeof(double) * A);
            if (j->done) j->done[(size_t)t * E + e] = ld;
            if (j->result) j->result[(size_t)t * E + e] = lr;
        }
    }
}

static void *thread_main(void *p) {
    const Job *j = (const Job *)p;
    if (j->T > 0) run_step_many(j); else run_step(j);
    return 0;
}

static void dispatch(Job base, int n_threads) {
    if (n_threads > 256) n_threads = 256;
    if (n_threads <= 1 || base.E < 2 * n_threads) { base.e0 = 0; base.e1 = base.E; thread_main(&base); return; }
    pthread_t th[256];
    Job jobs[256];
    for (int k = 0; k < n_threads; ++k) {
        jobs[k] = base;
        jobs[k].e0 = (int)((long long)base.E * k / n_threads);
        jobs[k].e1 = (int)((long long)base.E * (k + 1) / n_threads);
        pthread_create(&th[k], 0, thread_main, &jobs[k]);
    }
    for (int k = 0; k < n_threads; ++k) pthread_join(th[k], 0);
}

/* One step of E independent envs.  auto_reset != 0: envs that finish are reset in place and their
 * row of `obs` holds the first observation of the new episode (what the reference's caller stores,
 * train_fortattack.py:97-104); reward/done/result always describe the step just taken. */
int fa_oracle_step(int E, int n_guards, int n_attackers, int max_steps, double *st_f, uint8_t *st_i,
                   int32_t *time_step, const int32_t *actions, double *obs, double *rew,
                   uint8_t *done, uint8_t *result, double *margin, int auto_reset, uint64_t seed,
                   uint64_t env_id0, uint32_t *episode, int n_threads) {
    int A = n_guards + n_attackers;
    if (A > FA_MAX_AGENTS || A < 1) return -1;
    Job j = {0, E, 0, E, n_guards, A, max_steps, auto_reset, st_f, st_i, time_step, actions,
             obs, rew, done, result, margin, seed, env_id0, episode};
    dispatch(j, n_threads);
    return 0;
}

/* T consecutive steps with auto-reset on an action stream [T][E][A]; outputs are streams
 * obs [T][E][A][6], rew [T][E][A], done [T][E], result [T][E] (any may be NULL to skip storing). */
int fa_oracle_step_many(int T, int E, int n_guards, int n_attackers, int max_steps, double *st_f,
                        uint8_t *st_i, int32_t *time_step, const int32_t *actions, double *obs,
                        double *rew, uint8_t *done, uint8_t *result, uint64_t seed, uint64_t env_id0,
                        uint32_t *episode, int n_threads) {
    int A = n_guards + n_attackers;
    if (A > FA_MAX_AGENTS || A < 1 || T < 1) return -1;
    Job j = {T, E, 0, E, n_guards, A, max_steps, 1, st_f, st_i, time_step, actions,
             obs, rew, done, result, 0, seed, env_id0, episode};
    dispatch(j, n_threads);
    return 0;
}

/* Philox4x32-10 known-answer hook for tests. */
void fa_oracle_philox(uint32_t k0, uint32_t k1, uint32_t c[4]) { philox4x32_10(k0, k1, c); }
