"""ctypes binding of libisla_b200.so (the C ABI declared in include/isla_b200.h).

There is no CPU fallback: if the shared library is missing, cannot be loaded, or no CUDA device is
visible, every compute entry point raises.  ``load()`` itself works without a GPU so that the ABI
surface can be checked on a CPU-only box.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("ISLA_LIB_PATH") or os.path.join(_HERE, "lib", "libisla_b200.so")  # override: A/B experiments only

ISLA_OK = 0
ERR_INVALID, ERR_CUDA, ERR_CAPACITY, ERR_TRACE = -1, -2, -3, -4

ENV_CARTPOLE, ENV_PENDULUM, ENV_MOUNTAINCAR, ENV_FROZENLAKE, ENV_CURVE = 0, 1, 2, 3, 4
AGENT_VANILLA, AGENT_APW, AGENT_SPW, AGENT_DPW = 0, 1, 2, 3
BUDGET_TO_MAX_DEPTH, BUDGET_ZERO = 0, 1
EXPAND_RANDOM, EXPAND_VOO, EXPAND_GENETIC2, EXPAND_GENETIC = 0, 1, 2, 3
PRECISION_F32, PRECISION_F64 = 0, 1
KERNEL_AUTO, KERNEL_THREAD_PER_TREE, KERNEL_CTA_PER_TREE = 0, 1, 2


class IslaError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"isla_b200 error {status}: {message}")
        self.status = status


class IslaConfig(C.Structure):
    _fields_ = [
        ("env_id", C.c_int32), ("agent", C.c_int32), ("n_actions", C.c_int32), ("max_depth", C.c_int32),
        ("budget_mode", C.c_int32), ("rollouts_per_leaf", C.c_int32), ("expansion", C.c_int32),
        ("voo_n_samples", C.c_int32), ("voo_max_try", C.c_int32), ("precision", C.c_int32),
        ("kernel", C.c_int32), ("block_threads", C.c_int32), ("sim_capacity", C.c_int32), ("tuning", C.c_int32),
        ("action_grid0", C.c_int32), ("action_grid1", C.c_int32), ("cluster_ctas", C.c_int32), ("tree_parallel", C.c_int32),
        ("C", C.c_double), ("gamma", C.c_double), ("alpha1", C.c_double), ("k1", C.c_double),
        ("alpha2", C.c_double), ("k2", C.c_double), ("epsilon", C.c_double),
        ("seed", C.c_uint64), ("tree_id_base", C.c_uint64), ("n_trees", C.c_int64), ("virtual_value", C.c_double),
    ]


class IslaLayout(C.Structure):
    _fields_ = [
        ("kernel", C.c_int32), ("state_dim", C.c_int32), ("real_bytes", C.c_int32), ("n_actions", C.c_int32),
        ("action_dim", C.c_int32), ("reserved0", C.c_int32),
        ("slots_per_tree", C.c_int64), ("rec_offset", C.c_int64), ("rec_stride", C.c_int64),
        ("state_offset", C.c_int64), ("state_stride", C.c_int64), ("gfirst_offset", C.c_int64), ("gfirst_stride", C.c_int64),
        ("meta_offset", C.c_int64), ("path_offset", C.c_int64), ("pathg_offset", C.c_int64), ("pvown_offset", C.c_int64),
        ("counters_offset", C.c_int64), ("ln_offset", C.c_int64), ("ln_count", C.c_int64),
        ("disc_offset", C.c_int64), ("disc_count", C.c_int64), ("pw1_offset", C.c_int64), ("pw2_offset", C.c_int64), ("cumdisc_offset", C.c_int64),
        ("s_cap", C.c_int64), ("a_cap", C.c_int64),
        ("pool_offset", C.c_int64), ("pool_stride", C.c_int64), ("hash_offset", C.c_int64), ("hash_cap", C.c_int64),
        ("total_bytes", C.c_int64),
    ]


# every symbol include/isla_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "isla_last_error": (C.c_char_p, []),
    "isla_abi_version": (C.c_int, []),
    "isla_device_count": (C.c_int, []),
    "isla_engine_workspace_bytes": (C.c_int64, [C.POINTER(IslaConfig)]),
    "isla_engine_create": (C.c_int, [C.POINTER(IslaConfig), _P, C.c_int64, C.POINTER(_P)]),
    "isla_engine_destroy": (C.c_int, [_P]),
    "isla_engine_layout": (C.c_int, [_P, C.POINTER(IslaLayout)]),
    "isla_engine_set_roots": (C.c_int, [_P, _P, _P]),
    "isla_engine_set_rollout_heuristic": (C.c_int, [_P, _P, C.c_int32, C.c_int32]),
    "isla_engine_set_transition_noise": (C.c_int, [_P, _P, C.c_int32]),
    "isla_engine_run": (C.c_int, [_P, C.c_int32, _P]),
    "isla_engine_run_scripted": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, C.c_int64, _P]),
    "isla_engine_root_stats": (C.c_int, [_P, _P, _P, _P]),
    "isla_engine_best": (C.c_int, [_P, C.c_int32, _P, _P, _P]),
    "isla_engine_root_children": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P, _P]),
    "isla_engine_depth_stats": (C.c_int, [_P, C.c_int64, C.c_int32, _P, _P, _P, _P]),
    "isla_engine_pool_offsets": (C.c_int, [_P, _P]),
    "isla_engine_counters": (C.c_int, [_P, _P, _P]),
    "isla_engine_reseed": (C.c_int, [_P, C.c_uint64, C.c_uint64]),
    "isla_episode_step": (C.c_int, [C.c_int32, C.c_int32, C.c_int32, _P, _P, C.c_int64, C.c_int32, C.c_double, _P, _P, _P, _P]),
    "isla_merge_root_stats": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int32, C.c_int64, _P, _P, _P]),
    "isla_engine_root_merge": (C.c_int, [_P, C.c_int64, _P, _P]),
    "isla_kernel_launches": (C.c_int64, []),
    "isla_engine_reroot": (C.c_int, [_P, _P, _P, C.c_int32, _P]),
    "isla_env_step": (C.c_int, [C.c_int32, C.c_int32, _P, _P, C.c_int64, _P, _P, _P, _P]),
    "isla_rollout": (C.c_int, [C.c_int32, C.c_int32, _P, C.c_int64, C.c_int32, C.c_double, C.c_uint64, C.c_uint32,
                               C.c_int32, C.c_int32, _P, _P, _P]),
}

_lib = None


def load():
    """Load libisla_b200.so; raises (never falls back) if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise IslaError(ERR_CUDA, f"{LIB_PATH} not built: run `python -m islamcts_b200.build` "
                                      f"(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SYMBOLS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(status: int):
    if status < 0:
        raise IslaError(status, load().isla_last_error().decode("utf-8", "replace"))
    return status
