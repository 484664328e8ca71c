"""ctypes binding of libsmartcal_b200.so -- exactly the symbols declared in include/smartcal_b200.h.

There is no CPU fallback and no eager-PyTorch path: if the shared library is missing the import raises, and
creating an agent without a CUDA device raises (sc_agent_create returns SC_ERR_CUDA).
Build it with `python -c "import __graft_entry__ as g; g.build()"` (nvcc -gencode arch=compute_100a,code=sm_100a).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libsmartcal_b200.so")

SC_ABI_VERSION = 3
SC_FLAG_CRITIC_LOSS_SUM = 1   # include/smartcal_b200.h
SC_KIND_SAC, SC_KIND_TD3 = 0, 1
SC_CALIB_LOGSIGMA, SC_CALIB_NO_TRUNK, SC_CALIB_DEMIX_HINT, SC_CALIB_FLAT_META, SC_CALIB_SIDE32, SC_CALIB_TD3 = 1, 2, 4, 8, 16, 32
NET_ACTOR, NET_CRITIC_1, NET_CRITIC_2, NET_TARGET_CRITIC_1, NET_TARGET_CRITIC_2, NET_TARGET_ACTOR = range(6)
SLOT_PARAM, SLOT_ADAM_M, SLOT_ADAM_V = 0, 1, 2


class ScConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("kind", C.c_int32), ("state_dim", C.c_int32), ("n_actions", C.c_int32),
        ("batch_size", C.c_int32), ("mem_size", C.c_int32), ("prioritized", C.c_int32), ("use_hint", C.c_int32),
        ("update_actor_interval", C.c_int32), ("flags", C.c_int32),
        ("gamma", C.c_double), ("tau", C.c_double), ("lr_actor", C.c_double), ("lr_critic", C.c_double),
        ("reward_scale", C.c_double), ("alpha", C.c_double),
        ("max_action", C.c_float), ("per_clip", C.c_float), ("hint_threshold", C.c_float), ("admm_rho", C.c_float),
        ("rng_seed", C.c_uint64),
    ]


class ScTensorInfo(C.Structure):
    _fields_ = [("dev_ptr", C.c_void_p), ("rows", C.c_int64), ("cols", C.c_int64), ("pitch", C.c_int64),
                ("elem_bytes", C.c_int32), ("is_int", C.c_int32)]


class ScLearnStats(C.Structure):
    _fields_ = [("critic_1_loss", C.c_float), ("critic_2_loss", C.c_float), ("actor_loss", C.c_float),
                ("q1_mean", C.c_float), ("q2_mean", C.c_float), ("target_mean", C.c_float), ("logp_mean", C.c_float),
                ("gfun", C.c_float), ("rho", C.c_float),
                ("learn_counter", C.c_int32), ("did_learn", C.c_int32), ("actor_step", C.c_int32)]


class ScNoise(C.Structure):
    _fields_ = [("eps_next_host", C.c_void_p), ("eps_cur_host", C.c_void_p), ("eps_dual_host", C.c_void_p),
                ("td3_target_noise", C.c_double), ("has_td3_target_noise", C.c_int32), ("reserved", C.c_int32),
                ("per_u01_host", C.c_void_p), ("batch_idx_host", C.c_void_p)]


class ScCalibConfig(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("batch_size", C.c_int32), ("mem_size", C.c_int32), ("n_actions", C.c_int32),
                ("M", C.c_int32), ("img", C.c_int32), ("use_hint", C.c_int32), ("variant", C.c_int32),
                ("gamma", C.c_double), ("tau", C.c_double), ("lr_actor", C.c_double), ("lr_critic", C.c_double), ("alpha", C.c_double),
                ("hint_threshold", C.c_float), ("admm_rho", C.c_float), ("rng_seed", C.c_uint64),
                ("update_actor_interval", C.c_int32), ("reserved", C.c_int32)]


# name -> (restype, argtypes); every symbol of include/smartcal_b200.h
SIGNATURES = {
    "sc_last_error": (C.c_char_p, []),
    "sc_abi_version": (C.c_int, []),
    "sc_agent_arena_bytes": (C.c_size_t, [C.POINTER(ScConfig)]),
    "sc_agent_create": (C.c_int, [C.POINTER(ScConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "sc_agent_destroy": (C.c_int, [C.c_void_p]),
    "sc_agent_sync": (C.c_int, [C.c_void_p]),
    "sc_agent_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(ScTensorInfo)]),
    "sc_agent_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(ScTensorInfo)]),
    "sc_agent_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_double)]),
    "sc_agent_set_counters": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_double]),
    "sc_agent_set_scalars": (C.c_int, [C.c_void_p, C.c_float, C.c_float, C.c_float]),
    "sc_replay_store_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double]),
    "sc_replay_store_dev": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_double]),
    "sc_replay_fill_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sc_replay_set_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "sc_replay_sample_uniform_dev": (C.c_int, [C.c_void_p]),
    "sc_per_sample": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sc_per_update_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "sc_agent_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sc_agent_learn": (C.c_int, [C.c_void_p, C.POINTER(ScNoise), C.POINTER(ScLearnStats)]),
    "sc_agent_soft_update": (C.c_int, [C.c_void_p, C.c_float]),
    "sc_agent_params_changed": (C.c_int, [C.c_void_p]),
    "sc_agent_store_learn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_double, C.POINTER(ScNoise), C.POINTER(ScLearnStats)]),
    "sc_agent_set_stats_mode": (C.c_int, [C.c_void_p, C.c_int]),
    "sc_ensemble_create": (C.c_int, [C.POINTER(C.c_void_p), C.c_int, C.POINTER(C.c_void_p)]),
    "sc_ensemble_destroy": (C.c_int, [C.c_void_p]),
    "sc_ensemble_learn": (C.c_int, [C.c_void_p, C.POINTER(ScNoise), C.POINTER(ScLearnStats)]),
    "sc_ensemble_store_learn": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.POINTER(ScNoise), C.POINTER(ScLearnStats)]),
    "sc_ensemble_launch_count": (C.c_int64, [C.c_void_p]),
    "sc_agent_launch_count": (C.c_int64, [C.c_void_p]),
    "sc_agent_kernel_time": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_int64)]),
    "sc_agent_enable_kernel_timing": (C.c_int, [C.c_void_p, C.c_int]),
    # calibration conv-SAC group
    "sc_calib_arena_bytes": (C.c_size_t, [C.POINTER(ScCalibConfig)]),
    "sc_calib_create": (C.c_int, [C.POINTER(ScCalibConfig), C.c_void_p, C.c_size_t, C.c_void_p, C.POINTER(C.c_void_p)]),
    "sc_calib_destroy": (C.c_int, [C.c_void_p]),
    "sc_calib_sync": (C.c_int, [C.c_void_p]),
    "sc_calib_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_int, C.c_char_p, C.POINTER(ScTensorInfo)]),
    "sc_calib_buffer": (C.c_int, [C.c_void_p, C.c_char_p, C.POINTER(ScTensorInfo)]),
    "sc_calib_get_counters": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]),
    "sc_calib_set_counters": (C.c_int, [C.c_void_p, C.c_int64, C.c_int64]),
    "sc_calib_set_scalars": (C.c_int, [C.c_void_p, C.c_float, C.c_float]),
    "sc_calib_store_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int]),
    "sc_calib_set_batch_host": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int]),
    "sc_calib_act": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "sc_calib_learn": (C.c_int, [C.c_void_p, C.POINTER(ScNoise), C.POINTER(ScLearnStats)]),
    "sc_calib_soft_update": (C.c_int, [C.c_void_p, C.c_float]),
    "sc_calib_launch_count": (C.c_int64, [C.c_void_p]),
    "sc_calib_get_per_beta": (C.c_int, [C.c_void_p, C.POINTER(C.c_double)]),
    "sc_calib_set_per_beta": (C.c_int, [C.c_void_p, C.c_double]),
}

_lib = None


class SmartcalError(RuntimeError):
    pass


def load():
    """dlopen the CUDA library (once).  Raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is not built (run __graft_entry__.build()); "
            "smart-calibration_b200 has no CPU/PyTorch fallback")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)        # AttributeError if the .so does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    if lib.sc_abi_version() != SC_ABI_VERSION:
        raise ImportError(f"ABI mismatch: library {lib.sc_abi_version()} vs binding {SC_ABI_VERSION}")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise SmartcalError(f"libsmartcal_b200 error {rc}: {load().sc_last_error().decode()}")


def ptr(np_array):
    """host pointer of a C-contiguous numpy array (or None)"""
    return None if np_array is None else np_array.ctypes.data_as(C.c_void_p)
