"""ctypes binding of libffmp_b200.so — the C-ABI declared in include/ffmp_b200.h.

There is no CPU fallback: if the CUDA library is missing or the device is not a B200 every
entry point raises.  PyTorch is used by the callers only for device memory and streams.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FFMP_LIB_PATH") or os.path.join(HERE, "csrc", "libffmp_b200.so")   # override: A/B builds while tuning
ABI_VERSION = 1
COST_INF = 0x7FFFFFFF

EXPORTS = [
    "ffmp_last_error", "ffmp_abi_version", "ffmp_query_sizes", "ffmp_create", "ffmp_bind", "ffmp_destroy",
    "ffmp_reset", "ffmp_step", "ffmp_rollout", "ffmp_rollout_graphed", "ffmp_step_host", "ffmp_step_host_async", "ffmp_step_host_wait", "ffmp_obs_slot", "ffmp_set_obs_slot", "ffmp_set_terminal_obs", "ffmp_join", "ffmp_error_word", "ffmp_timing", "ffmp_launch_count", "ffmp_debug_trace", "ffmp_learner_input", "ffmp_scan", "ffmp_op_scan",
    "ffmp_feed_create", "ffmp_feed_handle", "ffmp_feed_connect", "ffmp_feed_info", "ffmp_feed_push", "ffmp_feed_wait", "ffmp_feed_release", "ffmp_feed_error", "ffmp_feed_destroy", "ffmp_pack_transitions",
    "ffmp_op_scenarios", "ffmp_op_flow_field_workspace", "ffmp_op_flow_field", "ffmp_op_rewarder", "ffmp_op_rewarder2", "ffmp_op_reward_calculator",
    "ffmp_replay_gather",
    "ffmp_qnet_create", "ffmp_qnet_load", "ffmp_qnet_forward", "ffmp_qnet_debug_activation", "ffmp_qnet_launch_count", "ffmp_qnet_destroy", "ffmp_qnet_last_error",
]


class NativeError(RuntimeError):
    pass


class Cfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_uint32), ("device", C.c_int32), ("num_envs", C.c_int32), ("grid", C.c_int32),
        ("window", C.c_int32), ("ring", C.c_int32), ("slots", C.c_int32), ("max_steps", C.c_int32),
        ("goal_mode", C.c_int32), ("block_shift", C.c_int32), ("p_thresh", C.c_uint32), ("env_id_base", C.c_uint32),
        ("seed", C.c_uint64), ("dt", C.c_float), ("regen_batch", C.c_uint32),
    ]


class Sizes(C.Structure):
    _fields_ = [(n, C.c_size_t) for n in
                ("cost", "flow", "scen", "state", "frames", "vec2", "vec1", "bytes1", "workspace")]


class Buffers(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in
                ("cost", "flow", "scen", "state", "frames", "rel_goal", "velocity", "reward", "done", "flags",
                 "term_rel_goal", "term_velocity", "fin_return", "fin_length", "workspace")]


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; fail loudly if it has not been built (no fallback path exists)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeError(
            f"{LIB_PATH} is missing: build it with `python -m flow_field_based_motion_planner_b200.build` "
            "(nvcc, sm_100a).  This package has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i32, u32, u64 = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64
    L.ffmp_last_error.restype = C.c_char_p
    L.ffmp_last_error.argtypes = []
    L.ffmp_abi_version.restype = u32
    L.ffmp_abi_version.argtypes = []
    L.ffmp_query_sizes.argtypes = [C.POINTER(Cfg), C.POINTER(Sizes)]
    L.ffmp_create.argtypes = [C.POINTER(Cfg), C.POINTER(vp)]
    L.ffmp_bind.argtypes = [vp, C.POINTER(Buffers)]
    L.ffmp_destroy.argtypes = [vp]
    L.ffmp_reset.argtypes = [vp, vp, vp]
    L.ffmp_step.argtypes = [vp, vp, vp]
    L.ffmp_rollout.argtypes = [vp, vp, i32, vp]
    L.ffmp_rollout_graphed.argtypes = [vp, vp, i32, vp]
    L.ffmp_step_host.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.ffmp_step_host_async.argtypes = [vp, vp, vp, vp, vp, vp, vp, vp]
    L.ffmp_step_host_wait.argtypes = [vp]
    L.ffmp_obs_slot.argtypes = [vp, C.POINTER(i32)]
    L.ffmp_set_obs_slot.argtypes = [vp, i32]
    L.ffmp_join.argtypes = [vp, vp]
    L.ffmp_replay_gather.argtypes = [i32, vp, C.c_size_t, i32, i32, i32, vp, vp, i32] + [vp] * 10
    L.ffmp_qnet_last_error.restype = C.c_char_p
    L.ffmp_qnet_create.argtypes = [i32, i32, C.POINTER(vp)]
    L.ffmp_qnet_load.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), vp]
    L.ffmp_qnet_forward.argtypes = [vp, i32, vp, i32, vp, vp, vp, i32, vp, vp]
    L.ffmp_qnet_debug_activation.argtypes = [vp, i32, i32, vp, C.POINTER(C.c_size_t), vp]
    L.ffmp_qnet_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.ffmp_qnet_destroy.argtypes = [vp]
    L.ffmp_set_terminal_obs.argtypes = [vp, vp]
    L.ffmp_timing.argtypes = [vp, i32, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(i32)]
    L.ffmp_launch_count.argtypes = [vp, C.POINTER(u64)]
    L.ffmp_debug_trace.argtypes = [vp, vp, vp]
    L.ffmp_learner_input.argtypes = [vp, vp, i32, C.c_float, vp]
    L.ffmp_scan.argtypes = [vp, i32, C.c_float, vp, vp, vp]
    L.ffmp_op_scan.argtypes = [i32, i32, i32, vp, i32, vp, i32, C.c_float, vp, vp, vp]
    L.ffmp_feed_create.argtypes = [i32, i32, i32, C.c_size_t, C.POINTER(vp)]
    L.ffmp_feed_handle.argtypes = [vp, vp]
    L.ffmp_feed_connect.argtypes = [vp, i32, vp]
    L.ffmp_feed_info.argtypes = [vp, C.POINTER(vp), C.POINTER(C.c_size_t), C.POINTER(C.c_size_t), C.POINTER(u32)]
    L.ffmp_feed_push.argtypes = [vp, vp, u32, C.c_double, vp]
    L.ffmp_feed_wait.argtypes = [vp, u32, C.c_double, C.POINTER(vp), vp]
    L.ffmp_feed_release.argtypes = [vp, u32, vp]
    L.ffmp_feed_error.argtypes = [vp, C.POINTER(u32), vp]
    L.ffmp_feed_destroy.argtypes = [vp]
    L.ffmp_pack_transitions.argtypes = [vp, vp, vp]
    L.ffmp_error_word.argtypes = [vp, C.POINTER(u32), vp]
    L.ffmp_op_scenarios.argtypes = [i32, i32, i32, u32, i32, i32, u64, vp, vp, vp, vp, vp]
    L.ffmp_op_flow_field_workspace.restype = C.c_size_t
    L.ffmp_op_flow_field_workspace.argtypes = [i32, i32]
    L.ffmp_op_flow_field.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp]
    L.ffmp_op_rewarder.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ffmp_op_rewarder2.argtypes = [i32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ffmp_op_reward_calculator.argtypes = [i32, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    for name in EXPORTS:
        fn = getattr(L, name)
        if fn.restype is C.c_int and name not in ("ffmp_last_error",):
            fn.restype = C.c_int
    if L.ffmp_abi_version() != ABI_VERSION:
        raise NativeError("libffmp_b200.so ABI version mismatch; rebuild the library")
    _lib = L
    return L


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().ffmp_last_error().decode("utf-8", "replace")
        raise NativeError(f"{what} failed (status {rc}): {msg}")
