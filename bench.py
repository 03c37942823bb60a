#!/usr/bin/env python
"""bench.py -- FortAttack step-only throughput (BASELINE.json config 2) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W

A "step" is ONE env.step() of every environment of the batch: 3 guards v 3 attackers, 4096 envs per
GPU (weak scaling: envs are independent, each rank owns its own shard, no collective on the data
path), uniform random actions, episode cap 100 with in-kernel auto-reset.

One JSON line (rank 0):
  value        agent-steps/s, whole job, device-resident: the K steps as ONE persistent launch (fa_step_many: state in
               registers, step t's actions read from HBM, step t's obs/reward/done/result written to HBM, auto-reset
               in the kernel), CUDA-event timed, max over ranks; `single_step_launches` gives the same K steps as K
               fa_step launches replayed from one CUDA graph (what a per-step env.step() costs on the device)
  e2e          the same metric through the host-buffer API (FortAttackBatch.step_many_host ->
               fa_step_many_host): every step's actions come from pinned host memory and every step's
               obs/reward/done/result are delivered to pinned host memory, the copies of neighbouring chunks
               of steps overlapped with the kernel; e2e.per_step_call = one synchronous fa_step_host call per
               step (what the numpy-facing env.step() costs)
  roofline     the step kernel: algorithmic bytes per launch (88 B/agent-step + 12 B/env-step, SURVEY 8d, x the E x K
               env-steps of the launch) / launch duration, against MEASURED_PEAKS.json hbm_gbs
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


def parity_sample(fab, torch, dev, steps=60):
    """Part of the cpu_baseline leg (the one place bench.py runs oracle/, as the CHECKER): the bench workload's float kernel
    teacher-forced against the float64 oracle for `steps` steps of E_PER_GPU envs -- every step starts from the oracle's own
    state.  Counts how many env-steps hold a laser decision whose float64 barycentric margin is below 1e-5 (the only place
    where a float-state predicate may legitimately decide differently) and how the masks compare inside and outside it."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import fa_oracle
    E = E_PER_GPU
    rng = np.random.RandomState(7)
    ora = fa_oracle.OracleEnv(E, NG, NA, max_steps=CAP, seed=3, n_threads=len(os.sched_getaffinity(0)))
    env = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=3, device=dev)
    ora.reset(); env.reset()
    near = differ_near = differ_far = contact_steps = 0
    worst = worst_contact = 0.0
    for _ in range(steps):
        act = rng.randint(0, 8, size=(E, A)).astype(np.int32)
        pos, alive = ora.st_f[:, :, 0:2].copy(), ora.st_i[:, :, 0] > 0
        dist = np.sqrt(((pos[:, :, None, :] - pos[:, None, :, :]) ** 2).sum(-1))
        touching = ((dist < 0.1) & alive[:, :, None] & alive[:, None, :] & ~np.eye(A, dtype=bool)[None]).any(axis=(1, 2))
        env.set_state(torch.from_numpy(ora.st_f.copy()), torch.from_numpy(ora.st_i.copy()), torch.from_numpy(ora.time_step.copy()),
                      torch.from_numpy(ora.episode.astype(np.int64)))
        obs, rew, done, res, margin = ora.step(act, auto_reset=True, want_margin=True)
        o, r, d, rs = env.step(torch.from_numpy(np.ascontiguousarray(act.T)).to(dev))
        o = o.cpu().numpy().swapaxes(0, 1)
        tight = margin < 1e-5
        bad = (d.cpu().numpy() != done) | (rs.cpu().numpy() != res) | (o[:, :, 0] != obs[:, :, 0]).any(axis=1)
        near += int(tight.sum()); differ_near += int((bad & tight).sum()); differ_far += int((bad & ~tight).sum())
        ok = ~tight & ~bad
        gate = np.abs(o - obs) / (1e-5 + 1e-5 * np.abs(obs))              # <= 1: inside the 1e-5 gate of the north star
        if (ok & ~touching).any():
            worst = max(worst, float(gate[ok & ~touching].max()))
        if (ok & touching).any():
            worst_contact = max(worst_contact, float(gate[ok & touching].max()))
        contact_steps += int(touching.sum())
    return {"env_steps": steps * E, "mapping": env.kernel_info()["mapping"], "margin": 1e-5,
            "env_steps_with_a_decision_inside_the_margin": near, "mask_mismatches_inside_the_margin": differ_near,
            "mask_mismatches_outside_the_margin": differ_far,
            "obs_error_over_gate_no_contact": worst, "obs_error_over_gate_in_contact": worst_contact,
            "env_steps_in_contact": contact_steps,
            "gate": "|got - ref| / (1e-5 + 1e-5 |ref|) <= 1; env-steps with two alive agents closer than 0.1 are listed separately: "
                    "the contact force amplifies the rounding of the float64 state to float32 by 1 / distance (tests/gpu_util.contact_slack)",
            "what": "float32 step kernel vs the float64 oracle, teacher-forced, alive / done / result masks compared per env-step; "
                    "double mode (FA_F64) has no such margin (tests)"}


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
    # the job's batch at N GPUs is N x 4096 envs (weak scaling): K steps of it = K * N * 4096 env.step() calls, bounded to
    # ~25 s at ~1.4e3 env-steps/s per core
    E_JOB = max(1, args.gpus) * E_PER_GPU
    budget = int(25.0 * 1400 * cores)
    k = max(1, min(args.steps, budget // E_JOB))
    ref = python_reference_rate(k * E_JOB, cores)
    if ref is not None:
        rate, dt, n = ref
        kind = "reference"
        sample = ("%d steps of the job's %d-env batch (%d GPU(s) x %d) = %d env.step() calls of the unmodified Python reference "
                  "(baseline/_ref/reference/gym_fortattack, numpy float64), one process per core x %d, %.1f s"
                  % (k, E_JOB, max(1, args.gpus), E_PER_GPU, n, cores, dt))
        what = ("the reference's own gym_fortattack env.step (byte copy of /root/reference under baseline/_ref, imported through "
                "stub modules for gym/pygame/pyglet), one Python process per host core, each stepping its own env")
    else:
        rate, dt, k, kind, sample = port_rate, port_dt, kp, "port", port["sample"]
        what = "oracle/fa_oracle.c (baseline/_ref not installed: python baseline/install_ref.py needs /root/reference)"
    line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
            "warmup": args.warmup, "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E_PER_GPU, "envs_total": E_JOB,
                       "note": "the CPU arm steps the whole job's batch (N x %d envs) on the host cores of rank 0's box" % E_PER_GPU,
                       "what_runs": what},
            "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample, "port": port},
            "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def bind_host_to_gpu(torch, local):
    """Place this rank's host side next to its GPU before any pinned allocation: (1) restrict the rank's threads to the CPUs
    NVML reports as NUMA-local to the GPU (when the container's cpuset has any), (2) ask the kernel to take this process's
    new pages -- the pinned buffers allocated afterwards -- from the GPU's own NUMA node (set_mempolicy MPOL_PREFERRED with
    the node sysfs names for the GPU's PCI function).  Returns a small report for the JSON line; never fatal."""
    rep = {"bound": False}
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(torch.cuda.get_device_properties(local).uuid)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinityWithinScope(h, (ncpu + 63) // 64, pynvml.NVML_AFFINITY_SCOPE_NODE)
        cpus = {i for i in range(ncpu) if (int(mask[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        rep.update({"gpu_numa_cpus": len(cpus), "allowed_cpus": len(allowed), "used_cpus": len(use)})
        if use and use != allowed:
            os.sched_setaffinity(0, use)
            rep["bound"] = True
        elif use:
            rep["note"] = "all allowed CPUs are already local to the GPU"
        else:
            rep["note"] = "none of the CPUs this container may use is local to the GPU"
        bus = pynvml.nvmlDeviceGetPciInfo(h).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()
        if len(bus.split(":")[0]) == 8:
            bus = bus[4:]
        with open("/sys/bus/pci/devices/%s/numa_node" % bus) as f:
            node = int(f.read().strip())
        rep["gpu_numa_node"] = node
        if node >= 0:
            import ctypes
            libc = ctypes.CDLL(None, use_errno=True)
            nodemask = (ctypes.c_ulong * 2)(0, 0)
            nodemask[node // 64] = 1 << (node % 64)
            rc = libc.syscall(238, 1, ctypes.byref(nodemask), 129)          # x86-64 set_mempolicy(MPOL_PREFERRED, mask, maxnode)
            rep["mempolicy"] = "preferred node %d" % node if rc == 0 else "set_mempolicy failed: errno %d" % ctypes.get_errno()
    except Exception as exc:
        rep["error"] = repr(exc)[:200]
    return rep


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
    binding = bind_host_to_gpu(torch, local)               # before any pinned allocation: first-touch places it NUMA-local
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        nodes = torch.tensor([binding.get("gpu_numa_node", -9), binding.get("used_cpus", -1),
                              1 if str(binding.get("mempolicy", "")).startswith("preferred") else 0], device=dev)
        allr = [torch.zeros_like(nodes) for _ in range(world)]
        dist.all_gather(allr, nodes)
        binding["by_rank"] = {"gpu_numa_node": [int(t[0]) for t in allr], "local_cpus_used": [int(t[1]) for t in allr],
                              "mempolicy_set": [int(t[2]) for t in allr]}
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
    # every step's obs / reward / done / result is stored: [K] slots for the persistent launch (0.69 MB per step)
    obs = torch.empty(K, A, E, 6, device=dev); rew = torch.empty(K, A, E, device=dev)
    done = torch.empty(K, E, dtype=torch.uint8, device=dev); res = torch.empty(K, E, dtype=torch.uint8, device=dev)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.25)
    reps = max(1, args.reps)

    def device_timed(fn):
        """CUDA-event time of fn()'s device work, best of `reps`, max over ranks.  A short spin kernel runs ahead of the first
        event so that everything fn() enqueues is already in the stream when the GPU reaches it (the events then bracket
        device execution, not this process's launch calls)."""
        best = None
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            barrier()
            torch.cuda._sleep(400000)                    # ~0.2 ms
            e0.record()
            fn()
            e1.record()
            barrier()
            ms = max_over_ranks(e0.elapsed_time(e1))
            best = ms if best is None else min(best, ms)
        return best

    # ---- headline: the K steps as ONE persistent launch (fa_step_many; the north-star's persistent kernel) ----------
    # state in registers for the whole launch, step t's actions read from HBM ([K][A][E] stream), step t's observations,
    # rewards, done and result flags written to HBM every step, auto-reset in the kernel
    persistent_launch = not args.single_step
    for _ in range(max(1, -(-W // K))):                   # >= W untimed steps
        env.step_many(acts, out=(obs, rew, done, res))
    torch.cuda.synchronize(dev)
    launches0 = env.launch_count()
    ms_many = device_timed(lambda: env.step_many(acts, out=(obs, rew, done, res)))
    launches_many = (env.launch_count() - launches0) // reps

    # ---- beside it: the same K steps as K single-step launches (fa_step, what env.step() costs) from one CUDA graph ---
    Ks = K if args.single_step else min(K, 512)
    R = min(Ks, 256)                                                                       # output ring

    def step(t):
        r = t % R
        env.step(acts[t], auto_reset=True, out=(obs[r], rew[r], done[r], res[r]))

    for t in range(min(W, Ks)):
        step(t)
    torch.cuda.synchronize(dev)
    graph = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for t in range(Ks):
                    step(t)
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()                                   # upload + one more untimed pass
        torch.cuda.synchronize(dev)
    ms_single = device_timed(graph.replay if graph is not None else (lambda: [step(t) for t in range(Ks)]))
    single = {"api": "fa_step (one launch per step)", "launch": "CUDA graph of %d fa_step launches" % Ks if graph is not None else "eager",
              "steps": Ks, "us_per_step": 1e3 * ms_single / Ks, "value": world * E * A * Ks / (ms_single * 1e-3)}
    if persistent_launch:
        ms_total, gpu_launches = ms_many, launches_many
    else:
        ms_total, gpu_launches = ms_single * K / Ks, K
    value = world * E * A * K / (ms_total * 1e-3)

    # ---- end to end: host actions -> device -> host results, every step -------------------------
    # ONE form of the call is the e2e figure: the chunk pipeline of fa_step_many_host (H2D copy | persistent step launch |
    # D2H copy on three streams), Ke steps of host actions in / host results out per call, called back to back over a
    # fixed window of >= 60 ms whatever --steps is (a 20-step region would be 0.3 ms of wall clock).  Beside it, in the
    # same run: the mapped single-launch form of the same call, and a copy-only probe (the same bytes per step moved
    # by cudaMemcpyAsync with no kernel at all) that shows what the host fabric allows on this box at this world size.
    Ke = args.e2e_steps
    hs = env.make_host_streams(Ke)
    g_e2e = torch.Generator().manual_seed(99 + rank)
    hs[0].copy_(torch.randint(0, 8, (Ke, A, E), generator=g_e2e, dtype=torch.int32))
    h_acts = hs[0]
    launches_e2e0 = env.launch_count()
    env.step_many_host(*hs)                                # warm-up: stream/event creation, staging buffer
    e2e_launches = env.launch_count() - launches_e2e0
    WINDOW = 0.06

    def timed_window(fn):
        """Calls fn() back to back for >= WINDOW seconds (all ranks run the same number of calls: rank 0's count after a
        calibration call is broadcast); returns seconds per call, max over ranks."""
        barrier()
        t0 = time.perf_counter(); fn(); one = time.perf_counter() - t0
        n = max(3, int(WINDOW / max(one, 1e-6)) + 1)
        if world > 1:
            tn = torch.tensor([n], device=dev); dist.broadcast(tn, 0); n = int(tn.item())
        barrier()
        t0 = time.perf_counter()
        for _ in range(n):
            fn()
        torch.cuda.synchronize(dev)
        dt = max_over_ranks(time.perf_counter() - t0)
        return dt / n, n, dt

    e2e_many_s, e2e_calls, e2e_window_s = timed_window(lambda: env.step_many_host(*hs))
    e2e_value = world * E * A * Ke / e2e_many_s
    e2e_form = "staged"
    env.step_many_host(*hs, chunk_steps=0)
    e2e_mapped_s, _, _ = timed_window(lambda: env.step_many_host(*hs, chunk_steps=0))
    e2e_mapped_many = world * E * A * Ke / e2e_mapped_s
    # copy-only probe: the bytes of Ke steps, H2D and D2H on two streams, no compute
    d_in = torch.empty(Ke, A, E, dtype=torch.int32, device=dev)
    d_out = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in hs[1:]]
    s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def copy_only():
        with torch.cuda.stream(s_in):
            d_in.copy_(h_acts, non_blocking=True)
        with torch.cuda.stream(s_out):
            for h_t, d_t in zip(hs[1:], d_out):
                h_t.copy_(d_t, non_blocking=True)
        s_in.synchronize(); s_out.synchronize()
    copy_only()
    copy_s, _, _ = timed_window(copy_only)
    copy_probe_value = world * E * A * Ke / copy_s
    del d_in, d_out
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

    # ---- roofline of the dominant kernel + larger batches --------------------------------------------------------
    # achieved = ALGORITHMIC bytes per launch / launch duration.  Algorithmic bytes = SURVEY.md 8(d)'s per-unit figure (88 B per
    # agent-step + 12 B per env-step: state in and out, action in, obs / reward / flags out) x the units one launch processes
    # (E envs x K steps for the persistent launch).  The persistent kernel keeps the state in registers between steps, so
    # the bytes it really has to move are fewer (32 B per agent-step + 2 B per env-step + the state once per launch):
    # `own_bytes` / `frac_own_bytes` give that stricter figure beside the prescribed one.
    peak, peak_src = peaks()
    info = env.kernel_info()
    steps_per_launch = K if persistent_launch else 1
    bytes_launch = steps_per_launch * E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP)
    launch_us = 1e3 * ms_total / (1 if persistent_launch else K)
    achieved = bytes_launch / (launch_us * 1e-6) / 1e9
    kname = {"agent": "fa::fa_step_wide_kernel", "group": "fa::fa_step_group_kernel", "env": "fa::fa_step_kernel"}[info["mapping"]]
    kname += "<3,3,float,%s>" % ("true" if persistent_launch else "false")
    traffic, traffic_src = None, None
    try:                                                   # dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full`
        with open(os.path.join(ROOT, "profiles", "step_traffic.json")) as f:
            tr = json.load(f).get("%s E=%d T=%d" % (kname, E, steps_per_launch))
        if tr:
            traffic, traffic_src = tr["dram_bytes"], tr["source"]
    except Exception:
        pass
    own_bytes = E * (steps_per_launch * (A * 32 + 2) + A * 56 + 12) if persistent_launch else bytes_launch
    roofline = {"kernel": kname, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "bytes_per_launch": bytes_launch, "steps_per_launch": steps_per_launch, "launch_us": launch_us,
                "own_bytes": own_bytes, "frac_own_bytes": own_bytes / (launch_us * 1e-6) / 1e9 / peak,
                "regime": "latency-bound: %.1f MB per step is L2-resident and one step is one warp's dependent instruction chain; "
                          "roofline_sweep shows the same kernels at batches that reach the HBM bound"
                          % (E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / 1e6),
                "regs": info["regs"], "block": info["block"], "grid": info["grid"], "mapping": info["mapping"]}
    single["frac"] = E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / (single["us_per_step"] * 1e-6) / 1e9 / peak
    extra = {}
    if rank == 0 and world == 1 and not args.quick:
        extra["roofline_sweep"] = sweep(fab, torch, dev, peak)
        extra["persistent"] = persistent(fab, torch, dev, peak, E, min(K, 1000))
        extra["persistent_1000_steps"] = persistent(fab, torch, dev, peak, E, 1000)
        extra["persistent_thread_per_agent"] = persistent(fab, torch, dev, peak, E, min(K, 1000), "agent")
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
        if world > 1:                         # config 5 sharded over the ranks (at world == 1 it is in rollout_ensemble above)
            try:
                extra["rollout_ensemble"] = rollout_config5_share(fab, torch, dev, dist=dist, world=world, rank=rank)
            except Exception as exc:
                extra["rollout_ensemble"] = {"error": repr(exc)}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": E, "n_guards": NG, "n_attackers": NA,
                       "launch": ("ONE persistent fa_step_many launch for the K steps: state in registers, every step's actions read "
                                  "from and every step's obs/reward/done/result written to HBM" if persistent_launch else
                                  ("CUDA graph of K fa_step launches" if graph is not None else "eager fa_step launches")),
                       "l2": ("not flushed: every step reads its own slice of the %.0f MB action stream and writes its own slice of "
                              "the %.0f MB of outputs (nothing is re-read between steps or timed repetitions except the "
                              "%.1f MB env state, by construction)"
                              % (acts.numel() * 4 / 1e6, (obs.numel() + rew.numel()) * 4 / 1e6, E * A * 28 / 1e6)),
                       "sharding": "independent env shards per rank, no data-path collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": Ke, "ms_per_step": 1e3 * e2e_many_s / Ke, "calls_timed": e2e_calls, "window_s": e2e_window_s,
                    "copy_only_probe": {"value": copy_probe_value, "d2h_gbs_per_gpu": d2h * Ke / copy_s / 1e9,
                                        "what": "the same h2d + d2h bytes per step moved by cudaMemcpyAsync on two streams, no kernel: "
                                                "the host-fabric ceiling of this box at this world size"},
                    "frac_of_copy_ceiling": e2e_value / copy_probe_value, "host_binding": binding,
                    "api": "FortAttackBatch.step_many_host / fa_step_many_host: ONE call per Ke steps; every step's actions "
                           "come from pinned host memory and every step's obs/reward/done/result land in pinned host memory",
                    "path": {"staged": "chunks of steps: H2D copy | persistent fa_step_many launch | D2H copy on three streams",
                             "mapped": "one persistent fa_step_many launch reading the actions and writing the results "
                                       "through mapped pinned host memory (no copy calls)"}[e2e_form],
                    "form": e2e_form, "mapped_single_launch_value": e2e_mapped_many,
                    "staged_launches_per_call": int(e2e_launches),
                    "pcie_d2h_gbs_per_gpu": d2h * Ke / e2e_many_s / 1e9,
                    "per_step_call": {"value": e2e_per_step_call, "ms_per_step": 1e3 * e2e_step_s / Ke,
                                      "api": "FortAttackBatch.step_host / fa_step_host, one synchronous call per step",
                                      "path": "kernel reads actions / writes results through mapped pinned host memory "
                                              "(no DMA calls)",
                                      "staged_copy_path_value": e2e_staged}},
            "gpu_launches": int(gpu_launches), "roofline": roofline, "single_step_launches": single}
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
        try:
            line["cpu_baseline"]["parity_sample"] = parity_sample(fab, torch, dev)
        except Exception as exc:
            line["cpu_baseline"]["parity_sample"] = {"error": repr(exc)[:300]}
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
    tr = ro.BatchedTrainer(E, NG, NA, num_steps=T, max_episode_steps=CAP, device=dev, seed=0, graph_update=False)   # (eager update timed below)
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
    # ... and with the optimizer steps replayed from CUDA graphs (BatchedTrainer's default on CUDA) and the teams overlapped
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
    # the same graph-replayed update with the two teams one after the other (overlap_teams=False; what earlier rounds measured)
    tr2.overlap_teams = False
    tr2.update()
    torch.cuda.synchronize(dev)
    e0, e1 = ev(), ev()
    e0.record(); tr2.update(); e1.record()
    torch.cuda.synchronize(dev)
    update_graph_seq_ms = e0.elapsed_time(e1)
    tr2.overlap_teams = True
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
            "ppo_update_graph_sequential_teams_ms": update_graph_seq_ms,
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
            "collect": "one CUDA graph of T x (2 mp_policy_kernel + fa_step with the rollout bookkeeping fused in), replayed" if tr._graph is not None else "eager",
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
        # the joint two-team step all-reduces ONE buffer holding both teams' gradients (+ 5 loss / normaliser terms each)
        flat = torch.zeros(sum(sum(p.numel() for p in pol.parameters()) + 5 for pol in tr.policies), device=dev)
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
           "collectives_per_optimizer_step": 0.5 if world > 1 else 0, "allreduce_us": ar_us, "allreduce_bytes": nbytes,
           "allreduce_share_of_update": (128 * ar_us * 1e-3 / upd) if ar_us else 0.0, "replica_weight_spread": spread,
           "update_path": "tcgen05 dense kernels + own attention / loss / Adam kernels, optimizer steps replayed from a CUDA graph "
                          "(NCCL all-reduce captured in it)",
           "teams_overlapped": True,
           "teams_note": "one rank: the two teams' updates (and recompute_old passes) run on two streams at the same time; several "
                         "ranks: the two teams' optimizer steps are the two branches of ONE captured graph around ONE all-reduce "
                         "of a flat buffer with both teams' gradients (128 collectives per update instead of 256; never two "
                         "collectives of the communicator in flight)",
           "policy_weights": "marlsave/tmp_1/ep2520.pt" if models is not None else "random init",
           "losses": [[float(x) for x in v] for v in vals]}
    del tr
    return out


def rollout_config5_share(fab, torch, dev, E=4096, T=32, K=5, dist=None, world=1, rank=0):
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
    tr = ro.BatchedTrainer(E, 5, 5, num_steps=T, max_episode_steps=CAP, device=dev, seed=0, env_id0=rank * E, attacker_ensemble=sds)
    if real:
        tr.load_models(shipped[-1])                                      # guards start from ep2520 (--pretrained-guard)
    for _ in range(2):
        tr.collect(); tr.wrap_horizon(); tr.after_update()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); tr.collect(); e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1)
    for fz in tr.fused + tr.ensemble:
        fz.check_status()
    if world > 1:                            # envs are independent: the job's time is the slowest rank's, the table sums the shards
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.all_reduce(tr.ensemble_results)
        dist.all_reduce(tr.ensemble_sums)
    return {"workload": "FortAttack 5v5 guards vs an ensemble of %d attacker checkpoints, %d envs per GPU x %d GPU(s) (BASELINE.json "
                        "configs[4] = 32 768 envs on 8), T=%d rollout" % (K, E, world, T),
            "rollout_agent_steps_per_s": world * E * 10 * T / (ms * 1e-3), "us_per_rollout_step": ms * 1e3 / T,
            "policy_launches_per_step": 2, "note": "one mp_forward for the guards, one mp_forward_ensemble serving all %d checkpoints" % K,
            "checkpoints": "marlsave/tmp_1/ep{220,650,1240,1600,2520}.pt attackers, ep2520 guards" if real else "random init (baseline/_ref absent)",
            "ensemble_table": tr.ensemble_table().round(3).tolist()}


def sweep(fab, torch, dev, peak):
    """Single-step kernel (fa_step) over batch sizes from L2-resident to far larger than the L2, each mapping where it
    applies: n launches replayed from one CUDA graph (so that the figure is the kernel's, not the CPU launch path's),
    CUDA events around the replay, best of 3."""
    out = []
    for E, mapping in ((4096, "group"), (4096, "agent"), (4096, "env"), (16384, "group"), (16384, "agent"), (16384, "env"),
                       (65536, "group"), (65536, "agent"), (65536, "env"), (1 << 18, "env"), (1 << 20, "env"), (1 << 22, "env")):
        env = fab.FortAttackBatch(E, NG, NA, max_steps=CAP, seed=0, device=dev, mapping=mapping)
        env.reset()
        n = 12 if E >= (1 << 20) else 40
        acts = torch.randint(0, 8, (n, A, E), device=dev, dtype=torch.int32)
        ring = 2 if E >= (1 << 20) else 8
        o = [(torch.empty(A, E, 6, device=dev), torch.empty(A, E, device=dev),
              torch.empty(E, dtype=torch.uint8, device=dev), torch.empty(E, dtype=torch.uint8, device=dev)) for _ in range(ring)]
        for t in range(3):
            env.step(acts[t], out=o[t % ring])
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                for t in range(n):
                    env.step(acts[t], out=o[t % ring])
        torch.cuda.current_stream(dev).wait_stream(side)
        graph.replay()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        us = 1e3 * best / n
        gbs = E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / (us * 1e-6) / 1e9
        out.append({"envs": E, "mapping": mapping, "launch_us": us, "agent_steps_per_s": E * A / (us * 1e-6), "achieved_gbs": gbs,
                    "frac": gbs / peak, "working_set_mb": E * (A * BYTES_AGENT_STEP + BYTES_ENV_STEP) / 1e6})
        del graph, env, acts, o
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
        torch.cuda._sleep(400000)                        # the launch is in the stream before the GPU reaches the first event
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
    ap.add_argument("--e2e-steps", type=int, default=256, help="steps per fa_step_many_host call of the e2e measurement")
    ap.add_argument("--no-graph", action="store_true", help="launch the single steps eagerly instead of from a CUDA graph")
    ap.add_argument("--single-step", action="store_true", help="headline = K single-step fa_step launches instead of one persistent fa_step_many launch")
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
