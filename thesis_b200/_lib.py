"""ctypes binding of thesis_b200/_librbpf.so (the C ABI in include/rbpf_b200.h).

There is no CPU fallback: if the shared library is missing or fails to load,
importing the compute path raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RBPF_LIB", os.path.join(_HERE, "_librbpf.so"))   # RBPF_LIB: tuning variants built by build.py --variant

RBPF_OK = 0
RBPF_ERR_ARG, RBPF_ERR_CUDA, RBPF_ERR_POOL, RBPF_ERR_RESAMPLE, RBPF_ERR_WORLD = -1, -2, -3, -4, -5
MOTION_ABSOLUTE, MOTION_VELOCITY, MOTION_UNICYCLE = 0, 1, 2


class RbpfConfig(C.Structure):
    _fields_ = [
        ("n_particles", C.c_int32), ("n_beams", C.c_int32), ("n_samples", C.c_int32),
        ("world_tiles_x", C.c_int32), ("world_tiles_y", C.c_int32), ("pool_subtiles", C.c_uint32),
        ("device", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32), ("flags", C.c_int32),
        ("stream", C.c_uint64), ("seed", C.c_uint64),
    ]


class RbpfPeerView(C.Structure):
    _fields_ = [
        ("ipc", (C.c_ubyte * 64) * 9), ("ptr", C.c_uint64 * 9), ("pid", C.c_int64), ("parity", C.c_int32),
        ("device", C.c_int32), ("n_particles", C.c_int32), ("pool_subtiles", C.c_uint32), ("nsub", C.c_int32),
        ("reserved", C.c_int32),
    ]


class RbpfStats(C.Structure):
    _fields_ = [
        ("pool_subtiles", C.c_uint32), ("pool_in_use", C.c_uint32), ("cow_copies", C.c_uint64),
        ("fresh_allocs", C.c_uint64), ("cells_dropped", C.c_uint64), ("resamples", C.c_uint64),
        ("match_failed", C.c_uint64), ("shared_refs", C.c_uint64), ("total_refs", C.c_uint64),
        ("refcount_sum", C.c_uint64), ("match_visits", C.c_uint64), ("match_points", C.c_uint64),
        ("match_runs", C.c_uint64),
        ("match_evals", C.c_uint64), ("ndt_evals", C.c_uint64), ("ndt_accepted", C.c_uint64),
        ("match_failed_zero", C.c_uint64),
    ]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_H = C.c_void_p

# name -> (restype, argtypes); must list every symbol include/rbpf_b200.h declares
SIGNATURES = {
    "rbpf_create": (C.c_int, [C.POINTER(RbpfConfig), C.POINTER(_H)]),
    "rbpf_destroy": (C.c_int, [_H]),
    "rbpf_last_error": (C.c_char_p, [_H]),
    "rbpf_set_scan": (C.c_int, [_H, _dp, _dp, C.c_int32]),
    "rbpf_motion": (C.c_int, [_H, C.c_int32, _dp, C.c_double, _dp]),
    "rbpf_scan_match": (C.c_int, [_H]),
    "rbpf_scan_match_adj": (C.c_int, [_H, _dp, C.c_int32]),
    "rbpf_weight": (C.c_int, [_H, _dp]),
    "rbpf_weight_guesses": (C.c_int, [_H, _dp]),
    "rbpf_integrate": (C.c_int, [_H, C.c_int32]),
    "rbpf_resample": (C.c_int, [_H, _dp, _ip, _ip]),
    "rbpf_step": (C.c_int, [_H, _dp, _dp, C.c_int32]),
    "rbpf_timing_enable": (C.c_int, [_H, C.c_int32]),
    "rbpf_timing_read": (C.c_int, [_H, _dp, _ip]),
    "rbpf_get_poses": (C.c_int, [_H, _dp]),
    "rbpf_get_covs": (C.c_int, [_H, _dp]),
    "rbpf_get_weights": (C.c_int, [_H, _dp]),
    "rbpf_set_poses": (C.c_int, [_H, _dp]),
    "rbpf_set_covs": (C.c_int, [_H, _dp]),
    "rbpf_set_weights": (C.c_int, [_H, _dp]),
    "rbpf_get_match": (C.c_int, [_H, _dp, _dp, _dp, _ip, _ip]),
    "rbpf_get_match_refine": (C.c_int, [_H, _ip]),
    "rbpf_set_refine": (C.c_int, [_H, C.c_int32]),
    "rbpf_get_resample_cumsum": (C.c_int, [_H, _dp]),
    "rbpf_set_match": (C.c_int, [_H, _dp, _dp, _ip]),
    "rbpf_get_match_slice": (C.c_int, [_H, C.c_int32, _ip]),
    "rbpf_export_tile": (C.c_int, [_H, C.c_int32, C.c_int32, C.c_int32, _dp, _ip]),
    "rbpf_list_tiles": (C.c_int, [_H, C.c_int32, _ip, C.c_int32, _ip]),
    "rbpf_occupied_points": (C.c_int, [_H, C.c_int32, _dp, C.c_int64, C.POINTER(C.c_int64)]),
    "rbpf_checkpoint_write": (C.c_int, [_H, C.c_char_p]),
    "rbpf_checkpoint_read": (C.c_int, [_H, C.c_char_p]),
    "rbpf_clear_errors": (C.c_int, [_H]),
    "rbpf_snapshot": (C.c_int, [_H]),
    "rbpf_restore": (C.c_int, [_H]),
    "rbpf_stats": (C.c_int, [_H, C.POINTER(RbpfStats)]),
    "rbpf_match_phase_clocks": (C.c_int, [_H, C.POINTER(C.c_uint64)]),
    "rbpf_synchronize": (C.c_int, [_H]),
    "rbpf_rot_step": (C.c_double, []),
    "rbpf_rot_count": (C.c_int32, []),
    "rbpf_weights_device_ptr": (C.c_int, [_H, C.POINTER(C.c_uint64)]),
    "rbpf_resample_global": (C.c_int, [_H, C.c_uint64, C.c_int32, _dp, _ip, _ip]),
    "rbpf_migrate_count": (C.c_int, [_H, _ip, C.c_int32, _ip, C.POINTER(C.c_int64)]),
    "rbpf_migrate_pack": (C.c_int, [_H, C.c_uint64]),
    "rbpf_migrate_bytes": (C.c_int64, [_H, C.c_int32, C.c_int32]),
    "rbpf_resample_apply_local": (C.c_int, [_H]),
    "rbpf_resample_apply_local_deferred": (C.c_int, [_H, C.c_uint64]),
    "rbpf_migrate_unpack": (C.c_int, [_H, C.c_uint64, C.c_int32, C.c_int32, _ip, _ip, C.c_int32]),
    "rbpf_resample_commit": (C.c_int, [_H]),
    "rbpf_peer_export": (C.c_int, [_H, C.POINTER(RbpfPeerView)]),
    "rbpf_peer_attach": (C.c_int, [_H, C.c_int32, C.POINTER(RbpfPeerView)]),
    "rbpf_migrate_pull": (C.c_int, [_H]),
    "rbpf_migrate_pull_async": (C.c_int, [_H, C.c_uint64]),
}

_LIB = None


def load():
    """Load the CUDA library; raises (loudly) when it has not been built."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "thesis_b200: %s is missing -- run `python -m thesis_b200.build` (needs nvcc); "
                "there is no CPU fallback" % LIB_PATH)
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _LIB = lib
    return _LIB
