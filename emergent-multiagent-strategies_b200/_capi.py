"""ctypes binding of libfortattack_b200.so (include/fortattack.h).

The library is the product: if it is missing or was not built, importing this module raises --
there is no CPU or PyTorch fallback for the step path.
"""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# FORTATTACK_B200_LIB: load another build of the same library (the diagnostic build with the policy kernel's phase trace)
LIB_PATH = os.environ.get("FORTATTACK_B200_LIB") or os.path.join(HERE, "libfortattack_b200.so")
PROBE_PATH = os.path.join(HERE, "libfortattack_probe.so")
CSRC = os.path.join(HERE, "csrc")

FA_ABI_VERSION = 1
FA_F32, FA_F64 = 0, 1
FA_MAX_TEAM = 5
FA_MAP_AUTO, FA_MAP_ENV, FA_MAP_AGENT, FA_MAP_GROUP = 0, 1, 2, 3

# every symbol include/fortattack.h declares
SYMBOLS = ("fa_abi_version", "fa_last_error", "fa_workspace_bytes", "fa_create", "fa_destroy", "fa_reset",
           "fa_step", "fa_step_many", "fa_step_host", "fa_step_many_host", "fa_host_stage_bytes", "fa_host_layout", "fa_get_state", "fa_set_state", "fa_alive_counts",
           "fa_set_max_steps", "fa_set_alive_end_buffer", "fa_set_rollout_outputs", "fa_launch_count", "fa_kernel_info")


class FaConfig(ctypes.Structure):
    _fields_ = [("n_envs", ctypes.c_int32), ("n_guards", ctypes.c_int32), ("n_attackers", ctypes.c_int32),
                ("max_steps", ctypes.c_int32), ("scalar", ctypes.c_int32), ("device", ctypes.c_int32),
                ("mapping", ctypes.c_int32), ("reserved", ctypes.c_int32), ("seed", ctypes.c_uint64), ("env_id0", ctypes.c_uint64)]


class FaState(ctypes.Structure):
    _fields_ = [("d_st_f", ctypes.c_void_p), ("d_st_i", ctypes.c_void_p), ("d_time_step", ctypes.c_void_p),
                ("d_episode", ctypes.c_void_p)]


class RlSmallMatmul(ctypes.Structure):
    """include/fortattack_rollout.h RlSmallMatmul"""
    _fields_ = [("A", ctypes.c_void_p), ("B", ctypes.c_void_p), ("C", ctypes.c_void_p), ("M", ctypes.c_int32), ("N", ctypes.c_int32),
                ("K", ctypes.c_int32), ("lda", ctypes.c_int32), ("ldb", ctypes.c_int32), ("ldc", ctypes.c_int32),
                ("trans_a", ctypes.c_int32), ("trans_b", ctypes.c_int32), ("accumulate", ctypes.c_int32)]


class RlAttnOperand(ctypes.Structure):
    """include/fortattack_rollout.h RlAttnOperand."""
    _fields_ = [("ptr", ctypes.c_void_p), ("batch_stride", ctypes.c_int64), ("row_stride", ctypes.c_int64)]


class TgTensor(ctypes.Structure):
    """include/fortattack_train.h TgTensor."""
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("numel", ctypes.c_longlong)]


class FrConfig(ctypes.Structure):
    """include/fortattack_render.h FrConfig."""
    _fields_ = [("n_envs", ctypes.c_int32), ("n_guards", ctypes.c_int32), ("n_attackers", ctypes.c_int32),
                ("width", ctypes.c_int32), ("height", ctypes.c_int32), ("draw_dead", ctypes.c_int32),
                ("reserved0", ctypes.c_int32), ("reserved1", ctypes.c_int32)]


class FaError(RuntimeError):
    pass


def build(force=False, jobs=6):
    """Compile csrc/ with nvcc for sm_100a (make -C csrc).  Works without a GPU."""
    import subprocess
    if force:
        subprocess.check_call(["make", "-C", CSRC, "clean"], stdout=subprocess.DEVNULL)
    subprocess.check_call(["make", "-C", CSRC, "-j%d" % jobs], stdout=subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FaError("%s is missing: build it with `make -C %s -j6` (or __graft_entry__.build()); "
                      "the FortAttack step path has no fallback implementation" % (LIB_PATH, CSRC))
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, u64p = ctypes.c_void_p, ctypes.c_int32, ctypes.POINTER(ctypes.c_uint64)
    i32p = ctypes.POINTER(ctypes.c_int32)
    L.fa_abi_version.restype = ctypes.c_int
    L.fa_last_error.restype = ctypes.c_char_p
    L.fa_workspace_bytes.argtypes = [ctypes.POINTER(FaConfig), ctypes.POINTER(ctypes.c_size_t)]
    L.fa_create.argtypes = [ctypes.POINTER(FaConfig), vp, ctypes.POINTER(vp)]
    L.fa_destroy.argtypes = [vp]
    L.fa_reset.argtypes = [vp, vp, vp, vp]
    L.fa_step.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    L.fa_step_many.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp]
    L.fa_step_host.argtypes = [vp, vp, vp, vp, vp, vp, ctypes.c_int, vp]
    szp = ctypes.POINTER(ctypes.c_size_t)
    L.fa_step_many_host.argtypes = [vp, ctypes.c_int, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    L.fa_host_stage_bytes.argtypes = [vp, ctypes.c_int, szp]
    L.fa_host_layout.argtypes = [vp, szp, szp, szp, szp]
    L.fa_get_state.argtypes = [vp, ctypes.POINTER(FaState), vp]
    L.fa_set_state.argtypes = [vp, ctypes.POINTER(FaState), vp]
    L.fa_alive_counts.argtypes = [vp, vp, vp]
    L.fa_set_max_steps.argtypes = [vp, i32]
    L.fa_set_alive_end_buffer.argtypes = [vp, vp]
    L.fa_set_rollout_outputs.argtypes = [vp, vp, vp, vp]
    L.fa_launch_count.argtypes = [vp, u64p]
    L.fa_kernel_info.argtypes = [vp, i32p, i32p, i32p, i32p, i32p]
    # every other entry point of the library (include/fortattack_policy.h, fortattack_rollout.h, mape_world.h).  ALL
    # prototypes are declared here, once: a function called without argtypes gets its pointers truncated to C ints.
    i64, u32, u64, f32, f64 = ctypes.c_longlong, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_float, ctypes.c_double
    P = ctypes.POINTER(RlAttnOperand)
    L.mp_forward.argtypes = [vp, vp, vp, i32, i32, i32, i32, u64, u64, vp, u64] + [vp] * 7 + [vp, i32, vp, vp, vp, vp]
    L.mp_kernel_info.argtypes = [i32, i32, i32p, i32p, i32p, i32p, i32p]
    L.mp_forward_ensemble.argtypes = [vp, i32, vp, vp, i32, i32, i32, i32, u64, u64, vp, u64] + [vp] * 8
    L.rl_gae.argtypes = [vp, vp, vp, vp, vp, vp, i32, i32, i32, f64, f64, vp]
    L.rl_rollout_bookkeeping.argtypes = [vp] * 7 + [i32, i32, vp]
    L.rl_gather_minibatch.argtypes = [vp] + [i32] * 8 + [vp] * 18
    L.rl_ppo_loss.argtypes = [vp] * 9 + [i32, f32, f32, f32] + [vp] * 5
    L.rl_ppo_loss_logits.argtypes = [vp] * 9 + [i32, i32, f32, f32, f32] + [vp] * 7
    L.rl_ppo_loss_logits_scratch_floats.argtypes = []
    L.rl_ppo_loss_logits_scratch_floats.restype = ctypes.c_size_t
    L.rl_attn_forward.argtypes = [P, P, P, P, vp, i32, i32, i32, i32, f32, i32, vp]
    L.rl_attn_backward.argtypes = [P, P, P, P, vp, P, P, P, i32, i32, i32, i32, f32, vp]
    L.rl_attn_mix_forward.argtypes = [P, P, P, P, vp, i32, i32, i32, i32, f32, i32, vp]
    L.rl_attn_mix_backward.argtypes = [P, P, P, vp, P, P, P, i32, i32, i32, i32, f32, vp]
    L.rl_relu_bwd_colsum_blocks.argtypes = [i64, i32]
    L.rl_relu_bwd_colsum.argtypes = [vp, vp, vp, vp, i64, i32, vp]
    L.rl_small_matmul.argtypes = [ctypes.POINTER(RlSmallMatmul), i32, vp]
    L.rl_colsum_blocks.argtypes = [i64, i32]
    L.rl_colsum.argtypes = [vp, i64, i32, i32, vp, vp, vp]
    L.rl_relu_bwd_colsum_ld.argtypes = [vp, i32, vp, i32, vp, i32, vp, i64, i32, vp]
    L.tg_packed_bytes.argtypes = [i32, i32]
    L.tg_packed_bytes.restype = ctypes.c_size_t
    L.tg_pack_weight.argtypes = [vp, i32, i32, i32, i32, vp, vp]
    L.tg_linear.argtypes = [vp, i32, i64, i32, vp, i32, vp, i32, i32, vp, i32, vp, vp]
    L.tg_linear_res.argtypes = [vp, i32, i64, i32, vp, i32, vp, i32, vp, i32, vp, i32, vp, vp]
    L.tg_wgrad_scratch_bytes.argtypes = [i32, i32]
    L.tg_wgrad_scratch_bytes.restype = ctypes.c_size_t
    L.tg_wgrad.argtypes = [vp, i32, i32, vp, i32, i32, i64, vp, i32, i32, vp, vp, vp]
    L.tg_adam_step.argtypes = [ctypes.POINTER(TgTensor), i32, f32, f32, f32, f32, f32, vp, vp, vp, vp, vp]
    L.tg_kernel_info.argtypes = [i32, i32p, i32p, i32p]
    L.tg_debug_wgrad_desc.argtypes = [u32, u32]
    L.tg_debug_staged.argtypes = [i32]
    L.tg_debug_wgrad_rows.argtypes = [i32]
    L.tg_debug_tma_out.argtypes = [i32]
    L.tg_debug_wgrad_staged.argtypes = [i32]
    L.mw_step.argtypes = [vp, vp, vp, vp, vp]
    L.mw_scenario_obs_dim.argtypes = [i32, i32, i32, i32]
    L.mw_scenario_callbacks.argtypes = [vp, i32, i32, vp, vp, vp, i32, vp, vp]
    L.mw_scenario_reset.argtypes = [vp, f64, f64, f64, f64, u64, u32, vp, vp, vp, vp]
    L.fr_render.argtypes = [ctypes.POINTER(FrConfig), vp, vp, vp, vp, i32, vp, vp]
    for name in SYMBOLS:
        getattr(L, name)   # AttributeError here = the library does not match the header
    if L.fa_abi_version() != FA_ABI_VERSION:
        raise FaError("libfortattack_b200.so has ABI %d, binding expects %d" % (L.fa_abi_version(), FA_ABI_VERSION))
    _lib = L
    return L


_probe = None


def probe_lib():
    """libfortattack_probe.so (include/fortattack_probe.h): hardware probes of the tensor-core operand layout used by the
    tests and the profiling scripts; test infrastructure, separate from the product library."""
    global _probe
    if _probe is None:
        P = ctypes.CDLL(PROBE_PATH)
        vp, i32, u32 = ctypes.c_void_p, ctypes.c_int32, ctypes.c_uint32
        P.mp_probe_gemm.argtypes = [vp, vp, vp, i32, i32, u32, u32, u32, vp, vp]
        P.mp_probe_timing.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, vp]
        _probe = P
    return _probe


def check(rc):
    if rc != 0:
        raise FaError("libfortattack_b200: %s (code %d)" % (lib().fa_last_error().decode(), rc))
