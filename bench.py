#!/usr/bin/env python
"""bench.py -- FortAttack step-only throughput (BASELINE.json config 2) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is ONE env.step() of every environment of the batch: 3 guards v 3 attackers, 4096 envs per
GPU (weak scaling: envs are independent, each rank owns its own shard, no collective on the data
path), uniform random actions, episode cap 100 with in-kernel auto-reset.

One JSON line (rank 0):
  value        agent-steps/s, whole job, device-resident: K launches of the fused single-step kernel
               (fa_step) replayed from one CUDA graph, actions pre-generated in HBM, CUDA-event timed,
               max over ranks
  e2e          the same metric through the host-buffer API (FortAttackBatch.step_many_host ->
               fa_step_many_host): every step's actions come from pinned host memory and every step's
               obs/reward/done/result are delivered to pinned host memory, the copies of neighbouring chunks
               of steps overlapped with the kernel; e2e.per_step_call = one synchronous fa_step_host call per
               step (what the numpy-facing env.step() costs)
  roofline     fa_step_kernel: algorithmic bytes per launch (88 B/agent-step + 12 B/env-step)
               / average launch duration over the timed region, against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline the CPU oracle port (oracle/fa_oracle.c, float64, pthreads on all host cores) on a
               bounded sample of the same workload (rank 0, N=1 only)
--impl reference times that CPU implementation as the whole arm (the reference itself is Python and
does not exist on the GPU box; oracle/ is its C restatement, pinned to it by tests/golden).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

NG, NA, E_PER_GPU, CAP = 3, 3, 4096, 100
A = NG + NA
BYTES_AGENT_STEP, BYTES_ENV_STEP = 88, 12           # SURVEY.md 8(d), single-step API, fp32
METRIC = "agent-steps/sec, FortAttack 3v3 x 4096 envs per GPU, step-only"
UNIT = "agent-steps/s"
WORKLOAD = ("FortAttack 3v3 (BASELINE.json configs[1]), %d envs per GPU, step-only, uniform random actions, "
            "episode cap %d, in-kernel auto-reset" % (E_PER_GPU, CAP))


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.lines, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
def cpu_oracle_rate(n_steps, warm, threads, rank_seed=0):
    """agent-steps/s of the CPU oracle port on the bench workload (E_PER_GPU envs), all outputs stored."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import fa_oracle
    env = fa_oracle.OracleEnv(E_PER_GPU, NG, NA, max_steps=CAP, seed=0, n_threads=threads)
    env.reset()
    rng = np.random.RandomState(rank_seed)
    chunk = 50
    acts = rng.randint(0, 8, size=(chunk, E_PER_GPU, A)).astype(np.int32)
    out = env.alloc_out(chunk)
    done_steps = 0
    while done_steps < warm:
        env.step_many(acts[:min(chunk, warm - done_steps)], out=tuple(o[:min(chunk, warm - done_steps)] for o in out))
        done_steps += chunk
    t0 = time.perf_counter()
    left = n_steps
    while left > 0:
        n = min(chunk, left)
        env.step_many(acts[:n], out=tuple(o[:n] for o in out))
        left -= n
    dt = time.perf_counter() - t0
    return E_PER_GPU * A * n_steps / dt, dt


REF_DIR = os.path.join(ROOT, "baseline", "_ref")


def _pyref_worker(rank, n_env_steps, warm, barrier, q):
    """One process = one core stepping the UNMODIFIED reference env (baseline/_ref/reference, imported through the stub
    modules of ref_shim.py) with uniform random actions; resets on done as train_fortattack.py:97-100 does."""
    try:
        sys.stdout = open(os.devnull, "w")                  # the env prints on every episode end (fortattack.py:208-220)
        os.environ["FA_REFERENCE_DIR"] = os.path.join(REF_DIR, "reference")
        sys.path.insert(0, REF_DIR)
        import numpy as np
        import ref_shim
        env, _ = ref_shim.make_ref_env(NG, NA, CAP)
        np.random.seed(1000 + rank)
        env.reset()
        acts = np.random.randint(0, 8, size=(n_env_steps + warm, A))
        for t in range(warm):
            if env.step(acts[t])[2]:
                env.reset()
        barrier.wait()
        t0 = time.perf_counter()
        for t in range(warm, warm + n_env_steps):
            if env.step(acts[t])[2]:
                env.reset()
        q.put((rank, n_env_steps, time.perf_counter() - t0, None))
    except Exception as exc:                                 # pragma: no cover
        try:
            barrier.abort()
        except Exception:
            pass
        q.put((rank, 0, 0.0, repr(exc)))


def python_reference_rate(env_steps_total, cores, warm=100):
    """agent-steps/s of the reference's own Python env.step, one process per host core.
    Returns (rate, seconds, env_steps_done) or None when baseline/_ref is not installed."""
    if not os.path.isdir(os.path.join(REF_DIR, "reference")):
        return None
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    per = max(1, (env_steps_total + cores - 1) // cores)
    barrier, q = ctx.Barrier(cores), ctx.Queue()
    procs = [ctx.Process(target=_pyref_worker, args=(r, per, warm, barrier, q), daemon=True) for r in range(cores)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=600) for _ in procs]
    for pr in procs:
        pr.join(timeout=30)
    bad = [r for r in res if r[3]]
    if bad:
        raise RuntimeError("python reference worker failed: %s" % bad[0][3])
    dt = max(r[2] for r in res)
    n = sum(r[1] for r in res)
    return n * A / dt, dt, n


def run_reference(args):
    """The reference arm: the UNMODIFIED Python reference (baseline/_ref) on all host cores, K steps of the bench's
    E_PER_GPU-env batch split over one process per core (bounded to ~25 s); the C port of the same arithmetic
    (oracle/fa_oracle.c) is timed beside it and reported under cpu_baseline.port."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    probe, _ = cpu_oracle_rate(20, 5, cores)
    kp = max(1, min(args.steps, int(30.0 * probe / (E_PER_GPU * A))))
    port_rate, port_dt = cpu_oracle_rate(kp, args.warmup, cores)
    port = {"value": port_rate, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d env.step() calls of %d envs (oracle/fa_oracle.c float64, %d pthreads), %.3f s" % (kp, E_PER_GPU, cores, port_dt)}
    # K steps of the 4096-env batch = K*4096 env-steps; bounded to ~25 s at ~1.4e3 env-steps/s per core
    budget = int(25.0 * 1400 * cores)
    k = max(1, min(args.steps, budget // E_PER_GPU))
    ref = python_reference_rate(k * E_PER_GPU, cores)
    if ref is not None:
        rate, dt, n = ref
        kind = "reference"
        sample = ("%d steps of the %d-env batch = %d env.step() calls of the unmodified Python reference "
                  "(baseline/_ref/reference/gym_fortattack, numpy float64), one process per core x %d, %.1f s"
                  % (k, E_PER_GPU, n, cores, dt))
        what = ("the reference's own gym_fortattack env.step (byte copy of /root/reference under baseline/_ref, imported through "
                "stub modules for gym/pygame/pyglet), one Python process per host core, each stepping its own env")
    else:
        rate, dt, k, kind, sample = port_rate, port_dt, kp, "port", port["sample"]
        what = "oracle/fa_oracle.c (baseline/_ref not installed: python baseline/install_ref.py needs /root/reference)"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "note": "CPU arm runs one shard of %d envs on the host cores whatever N is" % E_PER_GPU,
                       "what_runs": what},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "port": port},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import fortattack_b200 as fab

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the step path has no CPU implementation (use --impl reference)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    K, W, E = args.steps, max(3, args.warmup), args.envs

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    env = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=0, env_id0=rank * E, device=dev)
    env.reset()
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    acts = torch.randint(0, 8, (K, A, E), generator=g, device=dev, dtype=torch.int32)     # K*98 KB
    R = min(K, 256)                                                                        # output ring
    obs = torch.empty(R, A, E, 6, device=dev); rew = torch.empty(R, A, E, device=dev)
    done = torch.empty(R, E, dtype=torch.uint8, device=dev); res = torch.empty(R, E, dtype=torch.uint8, device=dev)

    def step(t):
        r = t % R
        env.step(acts[t], auto_reset=True, out=(obs[r], rew[r], done[r], res[r]))

    # ---- warm-up (eager), then capture the K-step rollout into one CUDA graph -------------------
    for t in range(W):
        step(t)
    torch.cuda.synchronize(dev)
    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for t in range(K):
                    step(t)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()                                   # upload + one more untimed pass
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    launches0 = env.launch_count()
    reps = max(1, args.reps)
    best = None
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        if graph is not None:
            graph.replay()
        else:
            for t in range(K):
                step(t)
        e1.record()
        barrier()
        ms = max_over_ranks(e0.elapsed_time(e1))
        best = ms if best is None else min(best, ms)
    gpu_launches = K if graph is not None else (env.launch_count() - launches0) // reps
    ms_total = best
    value = world * E * A * K / (ms_total * 1e-3)

    # ---- end to end: host actions -> device -> host results, every step -------------------------
    # (1) the stream call: Ke steps of host actions in, Ke steps of host obs/reward/done/result out, one
    #     fa_step_many_host call (chunks of steps: H2D copy | persistent step launch | D2H copy, overlapped)
    Ke = min(K, args.e2e_steps)
    hs = env.make_host_streams(Ke)
    hs[0].copy_(acts[:Ke].cpu())
    h_acts = hs[0]
    launches_e2e0 = env.launch_count()
    env.step_many_host(*hs)                                # warm-up: stream/event creation, staging buffer
    e2e_launches = env.launch_count() - launches_e2e0
    e2e_many_s = None
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        env.step_many_host(*hs)
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_many_s = dt if e2e_many_s is None else min(e2e_many_s, dt)
    e2e_staged_many = world * E * A * Ke / e2e_many_s
    # (2) the same stream call without staging: one persistent launch working through mapped pinned memory
    env.step_many_host(*hs, chunk_steps=0)
    e2e_mapped_s = None
    for _ in range(3):
        barrier()
        t0 = time.perf_counter()
        env.step_many_host(*hs, chunk_steps=0)
        dt = max_over_ranks(time.perf_counter() - t0)
        e2e_mapped_s = dt if e2e_mapped_s is None else min(e2e_mapped_s, dt)
    e2e_mapped_many = world * E * A * Ke / e2e_mapped_s
    # the headline e2e figure is the faster of the two forms of the SAME call (both are bound by the PCIe link)
    e2e_form = "staged" if e2e_many_s <= e2e_mapped_s else "mapped"
    e2e_many_s = min(e2e_many_s, e2e_mapped_s)
    e2e_value = world * E * A * Ke / e2e_many_s
    # (3) one synchronous call per step (the numpy-facing env.step of the facade): fa_step_host
    hb = env.make_host_buffers()
    ptr0, stride = h_acts.data_ptr(), A * E * 4
    lib, h, stream = fab._capi.lib(), env._h, torch.cuda.current_stream(dev).cuda_stream
    for t in range(W):
        env.step_host(h_acts[t], *hb[1:])
    barrier()
    t0 = time.perf_counter()
    for t in range(Ke):
        rc = lib.fa_step_host(h, ptr0 + t * stride, hb[1].data_ptr(), hb[2].data_ptr(), hb[3].data_ptr(),
                              hb[4].data_ptr(), 1, stream)
        if rc:
            fab._capi.check(rc)
    torch.cuda.synchronize(dev)
    e2e_step_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    e2e_per_step_call = world * E * A * Ke / e2e_step_s
    h2d = A * E * 4
    d2h = A * E * 6 * 4 + A * E * 4 + E + E
    clocks = sampler.stop() if rank == 0 else None
    e2e_staged = None
    if rank == 0 and world == 1 and not args.quick:       # the copy-based variant of the per-step call, for comparison
        os.environ["FA_HOST_PATH"] = "staged"
        env2 = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=0, device=dev)
        os.environ.pop("FA_HOST_PATH")
        env2.reset()
        hb2 = env2.make_host_buffers()
        for t in range(W):
            env2.step_host(h_acts[t], *hb2[1:])
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for t in range(Ke):
            lib.fa_step_host(env2._h, ptr0 + t * stride, hb2[1].data_ptr(), hb2[2].data_ptr(), hb2[3].data_ptr(),
                             hb2[4].data_ptr(), 1, stream)
        torch.cuda.synchronize(dev)
        e2e_staged = E * A * Ke / (time.perf_counter() - t0)
        del env2
    del hs

    # ---- roofline of the dominant kernel + larger batches + the persistent T-step kernel ---------
    peak, peak_src = peaks()
    bytes_launch = E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP)
    achieved = bytes_launch / (ms_total * 1e-3 / K) / 1e9
    info = env.kernel_info()
    roofline = {"kernel": "fa::fa_step_wide_kernel<3,3,float,false>" if info["mapping"] == "agent" else "fa::fa_step_kernel<3,3,float,false>", "bound": "hbm", "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak,
                # dram__bytes_read.sum + dram__bytes_write.sum of one launch under `ncu --set full` (profiles/): at 4096 envs the
                # state and actions are read from DRAM once, the 1.4 MB of results are still in L2 when the launch ends
                "traffic": 834048 if (E == 4096 and info["mapping"] == "agent") else None,
                "traffic_source": "profiles/r1c_step_3v3_E4096_ncu_full.txt", "peak_source": peak_src,
                "bytes_per_launch": bytes_launch, "launch_us": 1e3 * ms_total / K,
                "regs": info["regs"], "block": info["block"], "grid": info["grid"], "mapping": info["mapping"]}
    extra = {}
    if rank == 0 and world == 1 and not args.quick:
        extra["roofline_sweep"] = sweep(fab, torch, dev, peak)
        extra["persistent"] = persistent(fab, torch, dev, peak, E, min(K, 1000))
        extra["persistent_thread_per_env"] = persistent(fab, torch, dev, peak, E, min(K, 1000), "env")
        try:
            extra["rollout"] = rollout_config3(fab, torch, dev)
        except Exception as exc:                      # never lose the headline line to the secondary workload
            extra["rollout"] = {"error": repr(exc)}
        try:
            extra["rollout_5v5"] = rollout_config4_share(fab, torch, dev)
        except Exception as exc:
            extra["rollout_5v5"] = {"error": repr(exc)}
        try:
            extra["rollout_ensemble"] = rollout_config5_share(fab, torch, dev)
        except Exception as exc:
            extra["rollout_ensemble"] = {"error": repr(exc)}
    del env
    if not args.quick:
        try:
            extra["train"] = train_config4_share(torch, dev, dist, world, rank)
        except Exception as exc:
            extra["train"] = {"error": repr(exc)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "n_guards": NG, "n_attackers": NA,
                       "launch": "CUDA graph of K fa_step launches" if graph is not None else "eager fa_step launches",
                       "l2": ("not flushed: the action stream (%.0f MB) and the obs/reward output ring (%.0f MB) are "
                              "sized against the 126 MB L2; the %.1f MB env state is re-read every step by construction"
                              % (acts.numel() * 4 / 1e6, (obs.numel() + rew.numel()) * 4 / 1e6, E * A * 28 / 1e6)),
                       "sharding": "independent env shards per rank, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": 1e3 * e2e_many_s / Ke,
                    "api": "FortAttackBatch.step_many_host / fa_step_many_host: ONE call per Ke steps; every step's actions "
                           "come from pinned host memory and every step's obs/reward/done/result land in pinned host memory",
                    "path": {"staged": "chunks of steps: H2D copy | persistent fa_step_many launch | D2H copy on three streams",
                             "mapped": "one persistent fa_step_many launch reading the actions and writing the results "
                                       "through mapped pinned host memory (no copy calls)"}[e2e_form],
                    "form": e2e_form, "staged_chunks_value": e2e_staged_many, "mapped_single_launch_value": e2e_mapped_many,
                    "staged_launches_per_call": int(e2e_launches),
                    "pcie_d2h_gbs": d2h * Ke / e2e_many_s / 1e9,
                    "per_step_call": {"value": e2e_per_step_call, "ms_per_step": 1e3 * e2e_step_s / Ke,
                                      "api": "FortAttackBatch.step_host / fa_step_host, one synchronous call per step",
                                      "path": "kernel reads actions / writes results through mapped pinned host memory "
                                              "(no DMA calls)",
                                      "staged_copy_path_value": e2e_staged}},
            "gpu_launches": int(gpu_launches), "roofline": roofline}
    line.update(extra)
    if rank == 0 and world == 1 and not args.quick:
        cores = len(os.sched_getaffinity(0))
        probe, _ = cpu_oracle_rate(20, 5, cores)
        n = max(20, int(8.0 * probe / (E_PER_GPU * A)))
        rate, dt = cpu_oracle_rate(n, 5, cores)
        port = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "%d env.step() calls of %d envs, oracle/fa_oracle.c (float64, %d pthreads), %.1f s"
                          % (n, E_PER_GPU, cores, dt)}
        ref = None
        try:
            ref = python_reference_rate(int(12.0 * 1400 * cores), cores)      # ~12 s of the unmodified Python reference
        except Exception as exc:
            port["python_reference_error"] = repr(exc)
        if ref is not None:
            line["cpu_baseline"] = {"value": ref[0], "unit": UNIT, "cores": cores, "kind": "reference",
                                    "sample": "%d env.step() calls (3v3, uniform random actions, cap %d) of the unmodified Python "
                                              "reference from baseline/_ref, one process per core x %d, %.1f s"
                                              % (ref[2], CAP, cores, ref[1]),
                                    "per_core": ref[0] / cores, "port": port}
        else:
            line["cpu_baseline"] = port
    if rank == 0:
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def shipped_models(torch, ckpt=2520):
    """state_dicts of a shipped reference checkpoint (baseline/_ref/reference/marlsave/tmp_1/ep<ckpt>.pt: 10 entries,
    guards first) or None when the reference install is absent."""
    path = os.path.join(REF_DIR, "reference", "marlsave", "tmp_1", "ep%d.pt" % ckpt)
    if not os.path.exists(path):
        return None
    return torch.load(path, map_location="cpu")["models"]


def rollout_config3(fab, torch, dev, E=16384, T=128):
    """BASELINE.json configs[2]: 3v3, 16384 envs, full PPO rollout with the MPNN policy (fused tcgen05 forward +
    fused step + fused GAE), and one JointPPO update (4 epochs x 32 minibatches, torch autograd).  Reported next to
    the headline line, never mixed into it."""
    import importlib
    ro = importlib.import_module("emergent-multiagent-strategies_b200.rollout")
    pk = importlib.import_module("emergent-multiagent-strategies_b200.policy_kernel")
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(E, NG, NA, num_steps=T, max_episode_steps=CAP, device=dev, seed=0)
    models = shipped_models(torch)                                          # weights of marlsave/tmp_1/ep2520.pt
    if models is not None:
        tr.load_models(models)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(2):                                                      # eager warm-up, then the graph-capturing pass
        tr.collect(); tr.wrap_horizon(); tr.after_update()
    torch.cuda.synchronize(dev)
    e0, e1, e2 = ev(), ev(), ev()
    e0.record(); tr.collect(); e1.record(); tr.wrap_horizon(); e2.record()
    torch.cuda.synchronize(dev)
    collect_ms, wrap_ms = e0.elapsed_time(e1), e1.elapsed_time(e2)
    tr.recompute_old(); tr.wrap_horizon()                                   # what train_once() does before the update
    # the policy kernel alone (guards' team), CUDA events over 20 launches
    R = tr.roll
    f = tr.fused[0]
    out = f.forward(R.obs[0, 0:NG], R.obs[0, NG:A])
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(20):
        f.forward(R.obs[0, 0:NG], R.obs[0, NG:A], out=out)
    e1.record()
    torch.cuda.synchronize(dev)
    pol_us = e0.elapsed_time(e1) / 20 * 1e3
    with torch.no_grad():
        own, opp = R.obs[0, 0:NG].reshape(-1, 6), R.obs[0, NG:A].reshape(-1, 6)
        for _ in range(2):
            tr.policies[0].act(own, None, opp)
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(5):
            tr.policies[0].act(own, None, opp)
        e1.record()
        torch.cuda.synchronize(dev)
    torch_us = e0.elapsed_time(e1) / 5 * 1e3
    e0, e1 = ev(), ev()
    e0.record(); vals = tr.update(); e1.record()
    torch.cuda.synchronize(dev)
    update_ms = e0.elapsed_time(e1)
    fused = importlib.import_module("emergent-multiagent-strategies_b200.rlcore.fused")
    # comparison only: the same update with its dense products on cuBLAS fp32 (what round 1 measured; fused.DENSE is the
    # checker switch of the tests).  No TF32 figures any more: single-pass TF32 misses the gradient gate by 2-3 orders.
    fused.DENSE = "cublas"
    try:
        e0, e1 = ev(), ev()
        e0.record(); tr.update(); e1.record()
        torch.cuda.synchronize(dev)
        update_cublas_ms = e0.elapsed_time(e1)
    finally:
        fused.DENSE = "tcgen05"
    # ... and with the optimizer step replayed from a CUDA graph (opt-in JointPPO(graph_update=True))
    tr2 = ro.BatchedTrainer(E, NG, NA, num_steps=T, max_episode_steps=CAP, device=dev, seed=0, graph_update=True)
    if models is not None:
        tr2.load_models(models)
    tr2.collect(); tr2.recompute_old(); tr2.wrap_horizon()
    tr2.update()                                                            # three eager steps, capture, replays
    torch.cuda.synchronize(dev)
    e0, e1 = ev(), ev()
    e0.record(); tr2.update(); e1.record()
    torch.cuda.synchronize(dev)
    update_graph_ms = e0.elapsed_time(e1)
    e0, e1 = ev(), ev()
    e0.record(); tr2.recompute_old(); e1.record()
    torch.cuda.synchronize(dev)
    recompute_ms = e0.elapsed_time(e1)
    del tr2
    fused.tg_check_status(dev)
    for fz in tr.fused:
        fz.check_status()
    flop_row = 2 * (64 * 64 * 2 + 3 * 3 * 128 * 128 + 2 * 128 * 128)       # tensor-core MACs x2 per (agent, env) row
    mp = {}
    try:
        mp = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    tpeak = float(mp.get("bf16_tflops", 1590.0))
    ach = NG * E * flop_row / (pol_us * 1e-6) / 1e12
    return {"workload": "FortAttack 3v3 (BASELINE.json configs[2]), %d envs, T=%d rollout with the MPNN policy + one JointPPO update" % (E, T),
            "rollout_agent_steps_per_s": E * A * T / (collect_ms * 1e-3), "collect_ms": collect_ms,
            "us_per_rollout_step": collect_ms * 1e3 / T, "wrap_horizon_ms": wrap_ms, "ppo_update_ms": update_ms, "ppo_update_graph_ms": update_graph_ms,
            "ppo_update_cublas_fp32_ms": update_cublas_ms, "recompute_old_ms": recompute_ms,
            "policy_weights": "marlsave/tmp_1/ep2520.pt (baseline/_ref)" if models is not None else "random init (baseline/_ref absent)",
            "ppo_update": "4 epochs x 32 minibatches x 2 teams = 256 optimizer steps, fp32-grade arithmetic: every dense product on "
                          "this repo's tcgen05 kernels (tg_linear: fp16 hi/lo split, 3 MMAs per product; tg_wgrad: bf16 3-term split, "
                          "6 MMAs), folded message rounds with hand-written backward (rlcore/fused.py), attention forward/backward, "
                          "ReLU-backward + bias gradient, minibatch gather, clipped-PPO loss, clip + Adam (tg_adam_step) all own "
                          "kernels; ppo_update_cublas_fp32_ms = the same update with the products on cuBLAS fp32, for comparison; "
                          "recompute_old_ms = re-evaluating the rollout's log-probs / values with the update's forward "
                          "(BatchedTrainer.recompute_old, default in train_once)",
            "losses": vals,
            "policy_kernel": {"kernel": "mp::mp_policy_kernel", "us_per_team_forward": pol_us, "rows": NG * E,
                              "torch_module_act_us": torch_us, "speedup_vs_torch_module": torch_us / pol_us,
                              "roofline": {"bound": "tensor", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                                           "frac": ach / tpeak, "flop_per_row": flop_row,
                                           "peak_source": "MEASURED_PEAKS.json bf16_tflops (fp16 runs at the bf16 rate)" if mp else "fallback 1590"},
                              "info": f.kernel_info()},
            "collect": "one CUDA graph of T x (2 mp_policy_kernel + fa_step + bookkeeping), replayed" if tr._graph is not None else "eager",
            "gpu_launches_per_collect": {"mp_policy_kernel": 2 * T, "fa_step": T}}


def rollout_config4_share(fab, torch, dev, E=8192, T=32):
    """One GPU's share of BASELINE.json configs[3] (5v5, 65536 envs over 8 GPUs = 8192 envs per GPU): graph-replayed
    rollout collection with the MPNN policy; the gradient all-reduce of the update is exercised by tests/dist_train_gpu.py."""
    import importlib
    ro = importlib.import_module("emergent-multiagent-strategies_b200.rollout")
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(E, 5, 5, num_steps=T, max_episode_steps=CAP, device=dev, seed=0)
    for _ in range(2):
        tr.collect(); tr.wrap_horizon(); tr.after_update()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.collect(); e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    for fz in tr.fused:
        fz.check_status()
    return {"workload": "FortAttack 5v5, %d envs (one GPU's share of BASELINE.json configs[3]), T=%d rollout with the MPNN policy" % (E, T),
            "rollout_agent_steps_per_s": E * 10 * T / (ms * 1e-3), "us_per_rollout_step": ms * 1e3 / T,
            "envs_per_tile": tr.fused[0].kernel_info()["envs_per_tile"]}


def train_config4_share(torch, dev, dist, world, rank, E=8192, T=128):
    """BASELINE.json configs[3] (5v5, 65 536 envs over 8 GPUs = 8192 envs per GPU, PPO gradient all-reduce over NCCL): ONE full
    training iteration of this rank's share -- graph-replayed rollout collection, recompute_old, GAE, and the update's 256
    optimizer steps, each with ONE all-reduce of the flat gradient buffer (+ loss normaliser) when world > 1.  Runs on
    every rank; times are CUDA-event times, max over ranks."""
    import importlib
    ro = importlib.import_module("emergent-multiagent-strategies_b200.rollout")
    fused = importlib.import_module("emergent-multiagent-strategies_b200.rlcore.fused")
    pg = dist.group.WORLD if world > 1 else None
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(E, 5, 5, num_steps=T, max_episode_steps=CAP, device=dev, seed=0, env_id0=rank * E,
                           process_group=pg, graph_update=True)
    models = shipped_models(torch)
    if models is not None:
        tr.load_models(models)
    for _ in range(2):                       # eager pass, then the passes that capture the rollout and optimizer-step graphs
        tr.train_once()
    torch.cuda.synchronize(dev)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    ev[0].record(); tr.collect(); ev[1].record(); tr.recompute_old(); tr.wrap_horizon(); ev[2].record()
    vals = tr.update(); ev[3].record()
    torch.cuda.synchronize(dev)
    tr.after_update()
    t = torch.tensor([ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3]), ev[0].elapsed_time(ev[3])],
                     device=dev, dtype=torch.float64)
    ar_us, spread, nbytes = None, 0.0, 0
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        flat = torch.zeros(sum(p.numel() for p in tr.policies[0].parameters()) + 5, device=dev)
        nbytes = flat.numel() * 4
        for _ in range(5):
            dist.all_reduce(flat)
        torch.cuda.synchronize(dev)
        a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a0.record()
        for _ in range(50):
            dist.all_reduce(flat)
        a1.record()
        torch.cuda.synchronize(dev)
        ar = torch.tensor([a0.elapsed_time(a1) / 50 * 1e3], device=dev, dtype=torch.float64)
        dist.all_reduce(ar, op=dist.ReduceOp.MAX)
        ar_us = float(ar.item())
        w = torch.cat([p.detach().reshape(-1) for pol in tr.policies for p in pol.parameters()])
        lo, hi = w.clone(), w.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        spread = float((hi - lo).abs().max())
    for f in tr.fused:
        f.check_status()
    fused.tg_check_status(dev)
    for trn in tr.trainers:                  # the captured optimizer step holds NCCL kernels: drop it before the group goes away
        trn.release_graphs()
    tr._graph = None
    col, wrap, upd, tot = [float(x) for x in t.tolist()]
    out = {"workload": "FortAttack 5v5, %d envs per GPU x %d GPU(s) (BASELINE.json configs[3] = 65 536 envs on 8), T=%d: one training "
                       "iteration = rollout + recompute_old + GAE + JointPPO update (4 epochs x 32 minibatches x 2 teams)" % (E, world, T),
           "rollout_ms": col, "recompute_old_and_gae_ms": wrap, "update_ms": upd, "iteration_ms": tot,
           "agent_steps_per_s_trained": world * E * 10 * T / (tot * 1e-3), "optimizer_steps": 256,
           "collectives_per_optimizer_step": 1 if world > 1 else 0, "allreduce_us": ar_us, "allreduce_bytes": nbytes,
           "allreduce_share_of_update": (256 * ar_us * 1e-3 / upd) if ar_us else 0.0, "replica_weight_spread": spread,
           "update_path": "tcgen05 dense kernels + own attention / loss / Adam kernels, optimizer step replayed from a CUDA graph "
                          "(NCCL all-reduce captured in it)",
           "policy_weights": "marlsave/tmp_1/ep2520.pt" if models is not None else "random init",
           "losses": [[float(x) for x in v] for v in vals]}
    del tr
    return out


def rollout_config5_share(fab, torch, dev, E=4096, T=32, K=5):
    """One GPU's share of BASELINE.json configs[4] (guards-only training against an ensemble of 5 frozen attacker
    checkpoints, 32768 envs over 8 GPUs = 4096 envs per GPU, 5v5): rollout collection with a per-env, per-episode
    attacker draw = 1 guard forward + 1 ensemble forward (all K checkpoints in one launch) + 1 env step per rollout step (random-init checkpoints:
    the shipped checkpoints from baseline/_ref when present, else random-init stand-ins)."""
    import importlib
    ro = importlib.import_module("emergent-multiagent-strategies_b200.rollout")
    mp = importlib.import_module("emergent-multiagent-strategies_b200.mpnn")
    sds, ckpts = [], (220, 650, 1240, 1600, 2520)                       # arguments.py:62
    shipped = [shipped_models(torch, c) for c in ckpts[:K]]
    real = all(m is not None for m in shipped)
    for k in range(K):
        if real:
            sds.append(shipped[k][-1])                                   # attackers = models[-1] (learner.py:131-140)
        else:
            torch.manual_seed(100 + k)
            sds.append(mp.MPNN(action_space=ro._Shape(8), num_agents=5, num_opp_agents=5, input_size=6, hidden_dim=128).state_dict())
    torch.manual_seed(0)
    tr = ro.BatchedTrainer(E, 5, 5, num_steps=T, max_episode_steps=CAP, device=dev, seed=0, attacker_ensemble=sds)
    if real:
        tr.load_models(shipped[-1])                                      # guards start from ep2520 (--pretrained-guard)
    for _ in range(2):
        tr.collect(); tr.wrap_horizon(); tr.after_update()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.collect(); e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    for fz in tr.fused + tr.ensemble:
        fz.check_status()
    return {"workload": "FortAttack 5v5 guards vs an ensemble of %d attacker checkpoints, %d envs (one GPU's share of BASELINE.json "
                        "configs[4]), T=%d rollout" % (K, E, T),
            "rollout_agent_steps_per_s": E * 10 * T / (ms * 1e-3), "us_per_rollout_step": ms * 1e3 / T,
            "policy_launches_per_step": 2, "note": "one mp_forward for the guards, one mp_forward_ensemble serving all %d checkpoints" % K,
            "checkpoints": "marlsave/tmp_1/ep{220,650,1240,1600,2520}.pt attackers, ep2520 guards" if real else "random init (baseline/_ref absent)",
            "ensemble_table": tr.ensemble_table().round(3).tolist()}


def sweep(fab, torch, dev, peak):
    """Single-step kernel at batch sizes whose working set leaves the L2 (eager launches, CUDA events)."""
    out = []
    for E, mapping in ((4096, "agent"), (4096, "env"), (16384, "agent"), (16384, "env"), (65536, "agent"), (65536, "env"),
                       (1 << 20, "env"), (1 << 22, "env")):
        env = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=0, device=dev, mapping=mapping)
        env.reset()
        n = 12 if E >= (1 << 20) else 50
        acts = torch.randint(0, 8, (n, A, E), device=dev, dtype=torch.int32)
        o = (torch.empty(A, E, 6, device=dev), torch.empty(A, E, device=dev),
             torch.empty(E, dtype=torch.uint8, device=dev), torch.empty(E, dtype=torch.uint8, device=dev))
        for t in range(3):
            env.step(acts[t], out=o)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        for t in range(n):
            env.step(acts[t], out=o)
        e1.record()
        torch.cuda.synchronize(dev)
        us = 1e3 * e0.elapsed_time(e1) / n
        gbs = E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / (us * 1e-6) / 1e9
        out.append({"envs": E, "mapping": mapping, "launch_us": us, "agent_steps_per_s": E * A / (us * 1e-6), "achieved_gbs": gbs,
                    "frac": gbs / peak, "working_set_mb": E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / 1e6})
        del env, acts, o
    return out


def persistent(fab, torch, dev, peak, E, T, mapping="auto"):
    """fa_step_many: T steps in one launch, state in registers (32 B/agent-step + 2 B/env-step + state/T)."""
    env = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=0, device=dev, mapping=mapping)
    env.reset()
    acts = torch.randint(0, 8, (T, A, E), device=dev, dtype=torch.int32)
    out = (torch.empty(T, A, E, 6, device=dev), torch.empty(T, A, E, device=dev),
           torch.empty(T, E, dtype=torch.uint8, device=dev), torch.empty(T, E, dtype=torch.uint8, device=dev))
    env.step_many(acts, out=out)
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        e0.record()
        env.step_many(acts, out=out)
        e1.record()
        torch.cuda.synchronize(dev)
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    b = E * (A * 32 + 2 + (A * 56 + 12) / T)
    gbs = b / (best * 1e-3 / T) / 1e9
    return {"api": "fa_step_many", "mapping": env.kernel_info()["mapping"], "steps_per_launch": T, "envs": E, "us_per_step": 1e3 * best / T,
            "agent_steps_per_s": E * A * T / (best * 1e-3), "achieved_gbs": gbs, "frac": gbs / peak}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=E_PER_GPU, help="envs per GPU (default: the BASELINE config)")
    ap.add_argument("--reps", type=int, default=3, help="timed repetitions of the K-step region (best is reported)")
    ap.add_argument("--e2e-steps", type=int, default=500)
    ap.add_argument("--no-graph", action="store_true", help="launch the K steps eagerly instead of from a CUDA graph")
    ap.add_argument("--quick", action="store_true", help="skip the batch-size sweep, persistent kernel and CPU baseline")
    args = ap.parse_args()
    # The contract is ONE JSON line on stdout.  Libraries print there too (NCCL writes "NCCL version ..." to fd 1 when the
    # process group comes up), so fd 1 is pointed at stderr for the whole run and the line is written to the real stdout.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


if __name__ == "__main__":
    main()
