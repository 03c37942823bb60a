// fa_capi.cu -- the C ABI of include/fortattack.h: argument checking, workspace carving, kernel
// dispatch by (n_guards, n_attackers, scalar).  No device allocation happens here; every device byte
// belongs to the caller (a PyTorch tensor).
#include "../../include/fortattack.h"

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "fa_launch.h"

namespace fa {

__global__ void fa_alive_counts_kernel(const uint32_t *fl, int32_t *counts, int nE, int n_guards, int A) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= nE) return;
    int g = 0, a = 0;
    for (int i = 0; i < A; ++i) {
        const int al = fl[(size_t)i * nE + e] & F_ALIVE;
        if (i < n_guards) g += al; else a += al;
    }
    counts[e] = g;
    counts[nE + e] = a;
}

#define FA_EXTERN(NG)                                                                                                  \
    extern template cudaError_t launch_step_g<NG, float>(int, bool, int, const StepParams<float> &, int, int, cudaStream_t);  \
    extern template cudaError_t launch_step_g<NG, double>(int, bool, int, const StepParams<double> &, int, int, cudaStream_t); \
    extern template cudaError_t launch_reset_g<NG, float>(int, const StateView<float> &, const uint8_t *, float *, int,  \
                                                          uint64_t, uint64_t, int, int, cudaStream_t);                   \
    extern template cudaError_t launch_reset_g<NG, double>(int, const StateView<double> &, const uint8_t *, double *,    \
                                                           int, uint64_t, uint64_t, int, int, cudaStream_t);             \
    extern template cudaError_t step_attr_g<NG, float>(int, bool, int, cudaFuncAttributes *);                                 \
    extern template cudaError_t step_attr_g<NG, double>(int, bool, int, cudaFuncAttributes *);
FA_EXTERN(1) FA_EXTERN(2) FA_EXTERN(3) FA_EXTERN(4) FA_EXTERN(5)

#define FA_DISPATCH_NG(ng, CALL)          \
    switch (ng) {                         \
    case 1: return CALL(1);               \
    case 2: return CALL(2);               \
    case 3: return CALL(3);               \
    case 4: return CALL(4);               \
    case 5: return CALL(5);               \
    default: return cudaErrorInvalidValue; \
    }

template <typename R>
static cudaError_t launch_step(int ng, int na, bool many, int mapping, const StepParams<R> &p, int grid, int block,
                               cudaStream_t s) {
#define CALL(NG) launch_step_g<NG, R>(na, many, mapping, p, grid, block, s)
    FA_DISPATCH_NG(ng, CALL)
#undef CALL
}
template <typename R>
static cudaError_t launch_reset(int ng, int na, const StateView<R> &st, const uint8_t *mask, R *obs, int E, uint64_t seed,
                                uint64_t id0, int grid, int block, cudaStream_t s) {
#define CALL(NG) launch_reset_g<NG, R>(na, st, mask, obs, E, seed, id0, grid, block, s)
    FA_DISPATCH_NG(ng, CALL)
#undef CALL
}
template <typename R> static cudaError_t step_attr(int ng, int na, bool many, int mapping, cudaFuncAttributes *out) {
#define CALL(NG) step_attr_g<NG, R>(na, many, mapping, out)
    FA_DISPATCH_NG(ng, CALL)
#undef CALL
}

}  // namespace fa

// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";

// shared with the other translation units of the library (mp_policy.cu, rl_kernels.cu): records the calling
// thread's message for fa_last_error() and returns `code`
int fa_internal_fail(int code, const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}
#define fail fa_internal_fail
#define CUDA_TRY(expr)                                                                                     \
    do {                                                                                                   \
        cudaError_t e_ = (expr);                                                                           \
        if (e_ != cudaSuccess) return fail(FA_ECUDA, "%s: %s (%s)", #expr, cudaGetErrorString(e_), __func__); \
    } while (0)

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct FaHandle {
    FaConfig cfg;
    int A;
    size_t rs;          // sizeof(Real)
    char *ws;           // workspace base
    // state planes
    void *pv, *ap;
    uint32_t *fl;
    int32_t *tstep;
    uint32_t *episode;
    // device staging for fa_step_host
    int32_t *s_act;
    void *s_obs, *s_rew;
    uint8_t *s_done, *s_result;
    int block, grid;
    int pdl;            // launch fa_step with programmatic stream serialization (FA_PDL=1 enables)
    uint8_t *alive_end; // optional extra output of every step (fa_set_alive_end_buffer)
    int kmap;           // FA_MAP_ENV / FA_MAP_AGENT / FA_MAP_GROUP actually in use
    float *mask_next;   // optional rollout bookkeeping outputs of every step (fa_set_rollout_outputs)
    uint8_t *end_next;
    float *ep_rew;
    int sm_count;
    uint64_t launches;
    int host_path;      // fa_step_host: 0 auto, 1 staged only, 2 mapped only (env FA_HOST_PATH)
    // fa_step_many_host: copy streams and events of the chunk pipeline, created on first use
    cudaStream_t st_in, st_out;
    cudaEvent_t ev_in[2], ev_comp[2], ev_out[2];
    bool pipe_ready;
};

// Every entry point that launches, copies or touches the handle's streams/events runs on the handle's device whatever
// device the calling thread has current (a FortAttackBatch(device="cuda:1") used while cuda:0 is current), and restores it.
struct DevGuard {
    int prev = -1, dev;
    explicit DevGuard(const FaHandle *h) : dev(h->cfg.device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
    }
    ~DevGuard() {
        if (prev >= 0 && prev != dev) cudaSetDevice(prev);
    }
};

struct Carve {
    size_t off = 0;
    size_t take(size_t bytes) {
        size_t o = off;
        off = align_up(off + bytes, 256);
        return o;
    }
};

struct Layout {
    size_t pv, ap, fl, tstep, episode, s_act, s_obs, s_rew, s_done, s_result, total;
};

static Layout make_layout(const FaConfig &c) {
    const size_t E = (size_t)c.n_envs, A = (size_t)(c.n_guards + c.n_attackers);
    const size_t rs = c.scalar == FA_F64 ? 8 : 4;
    Carve k;
    Layout L;
    L.pv = k.take(A * E * 4 * rs);
    L.ap = k.take(A * E * 2 * rs);
    L.fl = k.take(A * E * 4);
    L.tstep = k.take(E * 4);
    L.episode = k.take(E * 4);
    L.s_act = k.take(A * E * 4);
    L.s_obs = k.take(A * E * 6 * rs);
    L.s_rew = k.take(A * E * rs);
    L.s_done = k.take(E);
    L.s_result = k.take(E);
    L.total = k.off;
    return L;
}

static int check_cfg(const FaConfig *c) {
    if (!c) return fail(FA_EINVAL, "config is NULL");
    if (c->n_envs < 1) return fail(FA_EINVAL, "n_envs must be >= 1 (got %d)", c->n_envs);
    if (c->n_guards < 1 || c->n_guards > FA_MAX_TEAM || c->n_attackers < 1 || c->n_attackers > FA_MAX_TEAM)
        return fail(FA_EINVAL, "unsupported team sizes %dv%d: kernels exist for 1..%d guards x 1..%d attackers",
                    c->n_guards, c->n_attackers, FA_MAX_TEAM, FA_MAX_TEAM);
    if (c->max_steps < 1) return fail(FA_EINVAL, "max_steps must be >= 1 (got %d)", c->max_steps);
    if (c->scalar != FA_F32 && c->scalar != FA_F64) return fail(FA_EINVAL, "scalar must be FA_F32 or FA_F64");
    if (c->mapping < FA_MAP_AUTO || c->mapping > FA_MAP_GROUP) return fail(FA_EINVAL, "mapping must be FA_MAP_AUTO/ENV/AGENT/GROUP");
    if ((uint64_t)c->n_envs * (uint64_t)(c->n_guards + c->n_attackers) * 6ull >= (1ull << 40))
        return fail(FA_EINVAL, "n_envs too large");
    return FA_OK;
}

template <typename R> static fa::StateView<R> view(const FaHandle *h) {
    fa::StateView<R> v;
    v.pv = reinterpret_cast<typename fa::VecT<R>::T4 *>(h->pv);
    v.ap = reinterpret_cast<typename fa::VecT<R>::T2 *>(h->ap);
    v.fl = h->fl;
    v.tstep = h->tstep;
    v.episode = h->episode;
    return v;
}

// Few envs: one warp per block so the launch spreads over as many SMs as possible (E=4096 -> 128
// SMs busy); many envs: 128-thread blocks (4 warps share one 6/12 KB obs stage).
static void pick_launch(FaHandle *h) {
    const int E = h->cfg.n_envs;
    // FA_MAP_AUTO (measured crossovers in DESIGN.md / profiles/r2i_step_mappings.jsonl):
    //   sub-warp group per env   while its warps number at most 4 per SM sub-partition (3v3: E <= 9 472, 5v5: E <= 4 736 on a
    //                            B200): a step lasts one warp's instruction chain, and the group kernel's is the shortest
    //   thread per agent         up to the batch where thread-per-env puts two warps on every sub-partition (E < 37 888)
    //   thread per env           beyond: the bandwidth-bound regime
    const int Gw = h->A <= 2 ? 2 : (h->A <= 4 ? 4 : (h->A <= 8 ? 8 : 16));
    const long long group_warps = ((long long)E * Gw + 31) / 32;
    h->kmap = h->cfg.mapping != FA_MAP_AUTO ? h->cfg.mapping
              : (group_warps <= 4ll * 4 * h->sm_count ? FA_MAP_GROUP : (E < 2 * 4 * 32 * h->sm_count ? FA_MAP_AGENT : FA_MAP_ENV));
    if (h->kmap == FA_MAP_AGENT) {
        h->block = 32 * h->A;
        h->grid = (E + 31) / 32;
        return;
    }
    if (h->kmap == FA_MAP_GROUP) {
        const int G = h->A <= 2 ? 2 : (h->A <= 4 ? 4 : (h->A <= 8 ? 8 : 16));
        const int warps = (E + 32 / G - 1) / (32 / G);
        int bw = 4;                                        // warps per block: fewer while the blocks would not cover the SMs twice
        while (bw > 1 && (warps + bw - 1) / bw < 2 * h->sm_count) bw >>= 1;
        if (const char *ev = getenv("FA_GROUP_BLOCK_WARPS")) {           // experiments (profiles/step_mappings.py)
            const int v = atoi(ev);
            if (v == 1 || v == 2 || v == 4) bw = v;
        }
        h->block = 32 * bw;
        h->grid = (warps + bw - 1) / bw;
        return;
    }
    int block = 128;
    while (block > 32 && (E + block - 1) / block < 2 * h->sm_count) block >>= 1;
    h->block = block;
    h->grid = (E + block - 1) / block;
}

#define NEED_HANDLE(h) \
    if (!(h)) return fail(FA_EINVAL, "handle is NULL")

template <typename R>
static cudaError_t do_step(FaHandle *h, bool many, int T, const int32_t *act, void *obs, void *rew, uint8_t *done,
                           uint8_t *result, int auto_reset, cudaStream_t s, bool rollout) {
    const FaConfig &c = h->cfg;
    fa::StepParams<R> p;
    p.st = view<R>(h);
    p.act = act;
    p.obs = static_cast<R *>(obs);
    p.rew = static_cast<R *>(rew);
    p.done = done;
    p.result = result;
    p.E = c.n_envs;
    p.T = T;
    p.max_steps = c.max_steps;
    p.auto_reset = auto_reset;
    // 16-byte vector stores of obs need every plane (stride E*6*sizeof(R)) to keep 16-byte alignment
    p.obs_vec_ok = ((size_t)c.n_envs * 6 * sizeof(R)) % 16 == 0 && ((uintptr_t)obs % 16) == 0;
    p.seed = c.seed;
    p.env_id0 = c.env_id0;
    p.alive_end = h->alive_end;
    p.mask_next = rollout ? h->mask_next : nullptr;
    p.end_next = rollout ? h->end_next : nullptr;
    p.ep_rew = rollout ? h->ep_rew : nullptr;
    p.pdl = h->pdl && !many;     // single-step launches chain through programmatic dependent launch
    return fa::launch_step<R>(c.n_guards, c.n_attackers, many, h->kmap, p, h->grid, h->block, s);
}

static int step_common(FaHandle *h, bool many, int T, const int32_t *d_actions, void *d_obs, void *d_reward,
                       uint8_t *d_done, uint8_t *d_result, int auto_reset, void *stream, bool rollout = false) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!d_actions) return fail(FA_EINVAL, "actions is NULL");
    if (T < 1) return fail(FA_EINVAL, "T must be >= 1 (got %d)", T);
    if (d_obs && (uintptr_t)d_obs % (2 * h->rs)) return fail(FA_EALIGN, "obs must be aligned to %zu bytes", 2 * h->rs);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    cudaError_t e = h->cfg.scalar == FA_F64
                        ? do_step<double>(h, many, T, d_actions, d_obs, d_reward, d_done, d_result, auto_reset, s, rollout)
                        : do_step<float>(h, many, T, d_actions, d_obs, d_reward, d_done, d_result, auto_reset, s, rollout);
    CUDA_TRY(e);
    h->launches += 1;
    return FA_OK;
}

extern "C" {

int fa_abi_version(void) { return FA_ABI_VERSION; }

const char *fa_last_error(void) { return g_err; }

int fa_workspace_bytes(const FaConfig *cfg, size_t *out_bytes) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!out_bytes) return fail(FA_EINVAL, "out_bytes is NULL");
    *out_bytes = make_layout(*cfg).total;
    return FA_OK;
}

int fa_create(const FaConfig *cfg, void *d_workspace, FaHandle **out) {
    int rc = check_cfg(cfg);
    if (rc) return rc;
    if (!d_workspace || !out) return fail(FA_EINVAL, "workspace / out is NULL");
    if ((uintptr_t)d_workspace % 256) return fail(FA_EALIGN, "workspace must be 256-byte aligned");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < 1) {
        cudaGetLastError();
        return fail(FA_ENODEVICE, "no CUDA device is visible; this library has no CPU path");
    }
    if (cfg->device < 0 || cfg->device >= ndev) return fail(FA_EINVAL, "device %d out of range (0..%d)", cfg->device, ndev - 1);
    CUDA_TRY(cudaSetDevice(cfg->device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, cfg->device));
    if (prop.major != 10)
        return fail(FA_ENODEVICE, "device %d is sm_%d%d; this library only carries sm_100a (B200) code", cfg->device,
                    prop.major, prop.minor);
    FaHandle *h = new (std::nothrow) FaHandle();
    if (!h) return fail(FA_EINVAL, "out of host memory");
    h->cfg = *cfg;
    h->A = cfg->n_guards + cfg->n_attackers;
    h->rs = cfg->scalar == FA_F64 ? 8 : 4;
    h->ws = static_cast<char *>(d_workspace);
    const Layout L = make_layout(*cfg);
    h->pv = h->ws + L.pv;
    h->ap = h->ws + L.ap;
    h->fl = reinterpret_cast<uint32_t *>(h->ws + L.fl);
    h->tstep = reinterpret_cast<int32_t *>(h->ws + L.tstep);
    h->episode = reinterpret_cast<uint32_t *>(h->ws + L.episode);
    h->s_act = reinterpret_cast<int32_t *>(h->ws + L.s_act);
    h->s_obs = h->ws + L.s_obs;
    h->s_rew = h->ws + L.s_rew;
    h->s_done = reinterpret_cast<uint8_t *>(h->ws + L.s_done);
    h->s_result = reinterpret_cast<uint8_t *>(h->ws + L.s_result);
    h->sm_count = prop.multiProcessorCount;
    h->launches = 0;
    h->host_path = 0;
    // Programmatic dependent launch of consecutive steps is opt-in (FA_PDL=1): measured on B200 at 3v3 x 4096 it gains
    // 2 % for eager launches (13.3 vs 13.6 us/step, CPU-launch-bound) and LOSES inside a replayed CUDA graph
    // (5.7 vs 3.7 us/step).
    h->pdl = 0;
    h->alive_end = nullptr;
    h->mask_next = nullptr;
    h->end_next = nullptr;
    h->ep_rew = nullptr;
    h->pipe_ready = false;
    if (const char *ev = getenv("FA_PDL")) h->pdl = ev[0] == '1';
    if (const char *hp = getenv("FA_HOST_PATH")) h->host_path = !strcmp(hp, "staged") ? 1 : (!strcmp(hp, "mapped") ? 2 : 0);
    pick_launch(h);
    const int E = cfg->n_envs, ib = 256, ig = (E + ib - 1) / ib;
    if (cfg->scalar == FA_F64) fa::fa_init_kernel<double><<<ig, ib>>>(view<double>(h), E, h->A);
    else fa::fa_init_kernel<float><<<ig, ib>>>(view<float>(h), E, h->A);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) {
        delete h;
        return fail(FA_ECUDA, "state initialisation failed: %s", cudaGetErrorString(e));
    }
    h->launches += 1;
    *out = h;
    return FA_OK;
}

int fa_destroy(FaHandle *h) {
    if (h && h->pipe_ready) {
        DevGuard dev_guard(h);
        cudaStreamDestroy(h->st_in);
        cudaStreamDestroy(h->st_out);
        for (int b = 0; b < 2; ++b) {
            cudaEventDestroy(h->ev_in[b]);
            cudaEventDestroy(h->ev_comp[b]);
            cudaEventDestroy(h->ev_out[b]);
        }
    }
    delete h;
    return FA_OK;
}

int fa_reset(FaHandle *h, const uint8_t *d_env_mask, void *d_obs, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const FaConfig &c = h->cfg;
    // fa_reset_kernel is always one thread per env
    int rblock = 128;
    while (rblock > 32 && (c.n_envs + rblock - 1) / rblock < 2 * h->sm_count) rblock >>= 1;
    const int rgrid = (c.n_envs + rblock - 1) / rblock;
    cudaError_t e;
    if (c.scalar == FA_F64)
        e = fa::launch_reset<double>(c.n_guards, c.n_attackers, view<double>(h), d_env_mask, static_cast<double *>(d_obs),
                                     c.n_envs, c.seed, c.env_id0, rgrid, rblock, s);
    else
        e = fa::launch_reset<float>(c.n_guards, c.n_attackers, view<float>(h), d_env_mask, static_cast<float *>(d_obs),
                                    c.n_envs, c.seed, c.env_id0, rgrid, rblock, s);
    CUDA_TRY(e);
    h->launches += 1;
    return FA_OK;
}

int fa_step(FaHandle *h, const int32_t *d_actions, void *d_obs, void *d_reward, uint8_t *d_done, uint8_t *d_result,
            int auto_reset, void *stream) {
    return step_common(h, false, 1, d_actions, d_obs, d_reward, d_done, d_result, auto_reset, stream, true);
}

int fa_step_many(FaHandle *h, int T, const int32_t *d_actions, void *d_obs, void *d_reward, uint8_t *d_done,
                 uint8_t *d_result, void *stream) {
    return step_common(h, true, T, d_actions, d_obs, d_reward, d_done, d_result, 1, stream, true);
}

// Device-visible alias of a host pointer if it is page-locked (cudaHostAlloc / cudaHostRegister, e.g. a
// pinned PyTorch tensor), else NULL.
static void *mapped_alias(const void *p) {
    if (!p) return nullptr;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? a.devicePointer : nullptr;
}

int fa_step_host(FaHandle *h, const int32_t *h_actions, void *h_obs, void *h_reward, uint8_t *h_done,
                 uint8_t *h_result, int auto_reset, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!h_actions) return fail(FA_EINVAL, "actions is NULL");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const size_t E = (size_t)h->cfg.n_envs, A = (size_t)h->A;
    const size_t nb_obs = A * E * 6 * h->rs, nb_rew = A * E * h->rs;

    // Path 1 (zero-copy): every buffer is page-locked -> the kernel reads the actions and writes its
    // results straight through the mapped host addresses; one launch + one stream sync, no DMA calls.
    if (h->host_path != 1) {
        void *m_act = mapped_alias(h_actions);
        void *m_obs = mapped_alias(h_obs), *m_rew = mapped_alias(h_reward);
        void *m_done = mapped_alias(h_done), *m_res = mapped_alias(h_result);
        const bool all = m_act && (!h_obs || m_obs) && (!h_reward || m_rew) && (!h_done || m_done) && (!h_result || m_res);
        if (all) {
            int rc = step_common(h, false, 1, static_cast<const int32_t *>(m_act), m_obs, m_rew,
                                 static_cast<uint8_t *>(m_done), static_cast<uint8_t *>(m_res), auto_reset, stream);
            if (rc) return rc;
            CUDA_TRY(cudaStreamSynchronize(s));
            return FA_OK;
        }
        if (h->host_path == 2) return fail(FA_EINVAL, "FA_HOST_PATH=mapped but a host buffer is not page-locked");
    }

    // Path 2 (staged): async copies through the workspace staging area; obs|reward|done|result are
    // contiguous there, so host buffers laid out the same way come back in ONE device->host copy.
    CUDA_TRY(cudaMemcpyAsync(h->s_act, h_actions, A * E * 4, cudaMemcpyHostToDevice, s));
    int rc = step_common(h, false, 1, h->s_act, h_obs ? h->s_obs : nullptr, h_reward ? h->s_rew : nullptr,
                         h_done ? h->s_done : nullptr, h_result ? h->s_result : nullptr, auto_reset, stream);
    if (rc) return rc;
    const bool packed = h_obs && h_reward && h_done && h_result &&
                        (char *)h_reward == (char *)h_obs + ((char *)h->s_rew - (char *)h->s_obs) &&
                        (char *)h_done == (char *)h_obs + ((char *)h->s_done - (char *)h->s_obs) &&
                        (char *)h_result == (char *)h_obs + ((char *)h->s_result - (char *)h->s_obs);
    if (packed) {
        const size_t total = (size_t)((char *)h->s_result - (char *)h->s_obs) + E;
        CUDA_TRY(cudaMemcpyAsync(h_obs, h->s_obs, total, cudaMemcpyDeviceToHost, s));
    } else {
        if (h_obs) CUDA_TRY(cudaMemcpyAsync(h_obs, h->s_obs, nb_obs, cudaMemcpyDeviceToHost, s));
        if (h_reward) CUDA_TRY(cudaMemcpyAsync(h_reward, h->s_rew, nb_rew, cudaMemcpyDeviceToHost, s));
        if (h_done) CUDA_TRY(cudaMemcpyAsync(h_done, h->s_done, E, cudaMemcpyDeviceToHost, s));
        if (h_result) CUDA_TRY(cudaMemcpyAsync(h_result, h->s_result, E, cudaMemcpyDeviceToHost, s));
    }
    CUDA_TRY(cudaStreamSynchronize(s));
    return FA_OK;
}

// Per-step byte counts of the five streams fa_step_many moves, and the layout of one chunk buffer of c steps
// (each region 256-byte aligned): actions | obs | reward | done | result.
struct ChunkLayout {
    size_t nb_act, nb_obs, nb_rew, nb_flag;
    size_t o_act, o_obs, o_rew, o_done, o_result, total;
};
static ChunkLayout chunk_layout(const FaHandle *h, size_t c) {
    const size_t E = (size_t)h->cfg.n_envs, A = (size_t)h->A;
    ChunkLayout L;
    L.nb_act = A * E * 4;
    L.nb_obs = A * E * 6 * h->rs;
    L.nb_rew = A * E * h->rs;
    L.nb_flag = E;
    Carve k;
    L.o_act = k.take(c * L.nb_act);
    L.o_obs = k.take(c * L.nb_obs);
    L.o_rew = k.take(c * L.nb_rew);
    L.o_done = k.take(c * L.nb_flag);
    L.o_result = k.take(c * L.nb_flag);
    L.total = k.off;
    return L;
}

int fa_host_stage_bytes(const FaHandle *h, int chunk_steps, size_t *out_bytes) {
    NEED_HANDLE(h);
    if (chunk_steps < 1) return fail(FA_EINVAL, "chunk_steps must be >= 1 (got %d)", chunk_steps);
    if (!out_bytes) return fail(FA_EINVAL, "out_bytes is NULL");
    *out_bytes = 2 * chunk_layout(h, (size_t)chunk_steps).total;
    return FA_OK;
}

static int pipe_init(FaHandle *h) {
    if (h->pipe_ready) return FA_OK;
    CUDA_TRY(cudaStreamCreateWithFlags(&h->st_in, cudaStreamNonBlocking));
    CUDA_TRY(cudaStreamCreateWithFlags(&h->st_out, cudaStreamNonBlocking));
    for (int b = 0; b < 2; ++b) {
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_in[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_comp[b], cudaEventDisableTiming));
        CUDA_TRY(cudaEventCreateWithFlags(&h->ev_out[b], cudaEventDisableTiming));
    }
    h->pipe_ready = true;
    return FA_OK;
}

int fa_step_many_host(FaHandle *h, int T, const int32_t *h_actions, void *h_obs, void *h_reward, uint8_t *h_done,
                      uint8_t *h_result, void *d_stage, size_t stage_bytes, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!h_actions) return fail(FA_EINVAL, "actions is NULL");
    if (T < 1) return fail(FA_EINVAL, "T must be >= 1 (got %d)", T);
    cudaStream_t s = static_cast<cudaStream_t>(stream);

    // Form 1 (no staging buffer): every host buffer is page-locked and the ONE persistent launch of fa_step_many
    // reads the action stream and writes the result streams through the mapped host addresses.
    if (!d_stage) {
        void *m_act = mapped_alias(h_actions);
        void *m_obs = mapped_alias(h_obs), *m_rew = mapped_alias(h_reward);
        void *m_done = mapped_alias(h_done), *m_res = mapped_alias(h_result);
        if (!(m_act && (!h_obs || m_obs) && (!h_reward || m_rew) && (!h_done || m_done) && (!h_result || m_res)))
            return fail(FA_EINVAL, "fa_step_many_host without a staging buffer needs page-locked host buffers");
        int rc = step_common(h, true, T, static_cast<const int32_t *>(m_act), m_obs, m_rew, static_cast<uint8_t *>(m_done),
                             static_cast<uint8_t *>(m_res), 1, stream);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(s));
        return FA_OK;
    }

    // Form 2 (chunk pipeline): the T steps are cut into chunks of c steps; chunk k's actions travel host->device on
    // one copy stream while chunk k-1 runs (one persistent fa_step_many launch per chunk on the caller's stream) and
    // chunk k-2's results travel device->host on the other copy stream.  Two chunk buffers, three events each.
    if ((uintptr_t)d_stage % 256) return fail(FA_EALIGN, "staging buffer must be 256-byte aligned");
    size_t c = (size_t)T;
    while (c > 1 && 2 * chunk_layout(h, c).total > stage_bytes) c = (c + 1) / 2;
    const ChunkLayout L = chunk_layout(h, c);
    if (2 * L.total > stage_bytes)
        return fail(FA_EINVAL, "staging buffer of %zu bytes is smaller than two one-step chunks (%zu bytes)", stage_bytes,
                    2 * L.total);
    int rc = pipe_init(h);
    if (rc) return rc;
    uint8_t *const alive_end0 = h->alive_end;
    char *const base = static_cast<char *>(d_stage);
    const char *ha = reinterpret_cast<const char *>(h_actions);
    int k = 0;
    for (size_t t0 = 0; t0 < (size_t)T; t0 += c, ++k) {
        const int b = k & 1;
        const size_t n = (size_t)T - t0 < c ? (size_t)T - t0 : c;
        char *buf = base + (size_t)b * L.total;
        // actions of this chunk: buffer b was last read by the launch of chunk k-2
        CUDA_TRY(cudaStreamWaitEvent(h->st_in, h->ev_comp[b], 0));
        CUDA_TRY(cudaMemcpyAsync(buf + L.o_act, ha + t0 * L.nb_act, n * L.nb_act, cudaMemcpyHostToDevice, h->st_in));
        CUDA_TRY(cudaEventRecord(h->ev_in[b], h->st_in));
        // the launch: its result regions were last drained by the copies of chunk k-2
        CUDA_TRY(cudaStreamWaitEvent(s, h->ev_in[b], 0));
        CUDA_TRY(cudaStreamWaitEvent(s, h->ev_out[b], 0));
        if (alive_end0) h->alive_end = alive_end0 + t0 * L.nb_flag;
        rc = step_common(h, true, (int)n, reinterpret_cast<const int32_t *>(buf + L.o_act), h_obs ? buf + L.o_obs : nullptr,
                         h_reward ? buf + L.o_rew : nullptr, h_done ? reinterpret_cast<uint8_t *>(buf + L.o_done) : nullptr,
                         h_result ? reinterpret_cast<uint8_t *>(buf + L.o_result) : nullptr, 1, stream);
        h->alive_end = alive_end0;
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(h->ev_comp[b], s));
        // results of this chunk
        CUDA_TRY(cudaStreamWaitEvent(h->st_out, h->ev_comp[b], 0));
        if (h_obs)
            CUDA_TRY(cudaMemcpyAsync((char *)h_obs + t0 * L.nb_obs, buf + L.o_obs, n * L.nb_obs, cudaMemcpyDeviceToHost, h->st_out));
        if (h_reward)
            CUDA_TRY(cudaMemcpyAsync((char *)h_reward + t0 * L.nb_rew, buf + L.o_rew, n * L.nb_rew, cudaMemcpyDeviceToHost, h->st_out));
        if (h_done)
            CUDA_TRY(cudaMemcpyAsync(h_done + t0 * L.nb_flag, buf + L.o_done, n * L.nb_flag, cudaMemcpyDeviceToHost, h->st_out));
        if (h_result)
            CUDA_TRY(cudaMemcpyAsync(h_result + t0 * L.nb_flag, buf + L.o_result, n * L.nb_flag, cudaMemcpyDeviceToHost, h->st_out));
        CUDA_TRY(cudaEventRecord(h->ev_out[b], h->st_out));
    }
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_out[0], 0));
    CUDA_TRY(cudaStreamWaitEvent(s, h->ev_out[1], 0));
    CUDA_TRY(cudaStreamSynchronize(s));
    return FA_OK;
}

int fa_host_layout(const FaHandle *h, size_t *off_reward, size_t *off_done, size_t *off_result, size_t *total) {
    NEED_HANDLE(h);
    if (off_reward) *off_reward = (size_t)((char *)h->s_rew - (char *)h->s_obs);
    if (off_done) *off_done = (size_t)((char *)h->s_done - (char *)h->s_obs);
    if (off_result) *off_result = (size_t)((char *)h->s_result - (char *)h->s_obs);
    if (total) *total = (size_t)((char *)h->s_result - (char *)h->s_obs) + (size_t)h->cfg.n_envs;
    return FA_OK;
}

int fa_get_state(FaHandle *h, const FaState *out, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!out || !out->d_st_f || !out->d_st_i || !out->d_time_step || !out->d_episode)
        return fail(FA_EINVAL, "FaState has a NULL member");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int E = h->cfg.n_envs, b = 128, g = (E + b - 1) / b;
    if (h->cfg.scalar == FA_F64)
        fa::fa_get_state_kernel<double><<<g, b, 0, s>>>(view<double>(h), out->d_st_f, out->d_st_i, out->d_time_step,
                                                        out->d_episode, E, h->A);
    else
        fa::fa_get_state_kernel<float><<<g, b, 0, s>>>(view<float>(h), out->d_st_f, out->d_st_i, out->d_time_step,
                                                       out->d_episode, E, h->A);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return FA_OK;
}

int fa_set_state(FaHandle *h, const FaState *in, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!in || !in->d_st_f || !in->d_st_i || !in->d_time_step || !in->d_episode)
        return fail(FA_EINVAL, "FaState has a NULL member");
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int E = h->cfg.n_envs, b = 128, g = (E + b - 1) / b;
    if (h->cfg.scalar == FA_F64)
        fa::fa_set_state_kernel<double><<<g, b, 0, s>>>(view<double>(h), in->d_st_f, in->d_st_i, in->d_time_step,
                                                        in->d_episode, E, h->A);
    else
        fa::fa_set_state_kernel<float><<<g, b, 0, s>>>(view<float>(h), in->d_st_f, in->d_st_i, in->d_time_step,
                                                       in->d_episode, E, h->A);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return FA_OK;
}

int fa_alive_counts(FaHandle *h, int32_t *d_counts, void *stream) {
    NEED_HANDLE(h);
    DevGuard dev_guard(h);
    if (!d_counts) return fail(FA_EINVAL, "counts is NULL");
    const int E = h->cfg.n_envs, b = 128, g = (E + b - 1) / b;
    fa::fa_alive_counts_kernel<<<g, b, 0, static_cast<cudaStream_t>(stream)>>>(h->fl, d_counts, E, h->cfg.n_guards, h->A);
    CUDA_TRY(cudaGetLastError());
    h->launches += 1;
    return FA_OK;
}

int fa_set_max_steps(FaHandle *h, int32_t max_steps) {
    NEED_HANDLE(h);
    if (max_steps < 1) return fail(FA_EINVAL, "max_steps must be >= 1 (got %d)", max_steps);
    h->cfg.max_steps = max_steps;
    return FA_OK;
}

int fa_set_alive_end_buffer(FaHandle *h, uint8_t *d_alive_end) {
    NEED_HANDLE(h);
    h->alive_end = d_alive_end;
    return FA_OK;
}

int fa_set_rollout_outputs(FaHandle *h, float *d_mask_next, uint8_t *d_end_next, float *d_ep_reward) {
    NEED_HANDLE(h);
    h->mask_next = d_mask_next;
    h->end_next = d_end_next;
    h->ep_rew = d_ep_reward;
    return FA_OK;
}

int fa_launch_count(const FaHandle *h, uint64_t *out) {
    NEED_HANDLE(h);
    if (!out) return fail(FA_EINVAL, "out is NULL");
    *out = h->launches;
    return FA_OK;
}

int fa_kernel_info(const FaHandle *h, int32_t *regs, int32_t *block, int32_t *grid, int32_t *smem, int32_t *mapping) {
    NEED_HANDLE(h);
    cudaFuncAttributes a;
    cudaError_t e = h->cfg.scalar == FA_F64
                        ? fa::step_attr<double>(h->cfg.n_guards, h->cfg.n_attackers, false, h->kmap, &a)
                        : fa::step_attr<float>(h->cfg.n_guards, h->cfg.n_attackers, false, h->kmap, &a);
    CUDA_TRY(e);
    if (regs) *regs = a.numRegs;
    if (block) *block = h->block;
    if (grid) *grid = h->grid;
    if (smem) *smem = (int32_t)a.sharedSizeBytes;
    if (mapping) *mapping = h->kmap;
    return FA_OK;
}

}  // extern "C"
