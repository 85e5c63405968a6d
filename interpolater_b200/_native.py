"""ctypes binding of the C ABI declared in include/interpolater_b200.h.

The shared library is built in-tree by ``__graft_entry__.build()`` (nvcc, sm_100a) and loaded from
``interpolater_b200/lib``.  There is no CPU fallback: a missing library or a missing sm_100 device
raises, loudly, on first use.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("IPR_LIB_OVERRIDE") or os.path.join(_HERE, "lib", "libinterpolater_b200.so")

IPR_OK = 0
IPR_ECANCELLED = -5
ERRLEN = 512
#: phases of ipr_plan_profile_read (include/interpolater_b200.h, IPR_PHASE_*)
N_PHASES = 32
PHASES = {"prep": 0, "weights": 1, "contract": 2, "slice": 3, "mask_gemm": 4, "fixup": 5, "post": 6}
PHASE_RADIUS0 = 8

#: every symbol include/interpolater_b200.h declares (checked by the CPU test-suite)
EXPORTED_SYMBOLS = [
    "ipr_version", "ipr_device_count", "ipr_na_real",
    "ipr_idw_f64", "ipr_cressman_f64", "ipr_idw_post_f64", "ipr_cressman_post_f64", "ipr_plan_set_post",
    "ipr_plan_create", "ipr_plan_destroy", "ipr_plan_set_obs", "ipr_plan_set_obs_device",
    "ipr_plan_idw", "ipr_plan_cressman", "ipr_plan_sync", "ipr_plan_reset_cache",
    "ipr_plan_timer_start", "ipr_plan_timer_stop", "ipr_plan_launch_count", "ipr_plan_zero_pairs",
    "ipr_plan_stage_blocks", "ipr_plan_profile", "ipr_plan_profile_read",
    "ipr_plan_stream", "ipr_host_alloc", "ipr_host_free", "ipr_release_cached_memory",
    "ipr_job_start_idw", "ipr_job_start_cressman", "ipr_job_wait", "ipr_job_cancel", "ipr_job_finish",
    "ipr_cell_split_bounds", "ipr_idw_kriging_f64", "ipr_plan_idw_kriging",
    "ipr_kriging_f64", "ipr_plan_kriging",
]


class NativeError(RuntimeError):
    """Raised when the CUDA engine reports an error (code + message from the C ABI)."""

    def __init__(self, code: int, message: str):
        super().__init__(f"interpolater_b200 native error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Load (once) and type the native library.  Raises if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). interpolater_b200 has no CPU fallback."
        )
    lib = C.CDLL(LIB_PATH)
    dp = C.POINTER(C.c_double)
    ip = C.POINTER(C.c_int)
    vp = C.c_void_p
    lib.ipr_version.restype = C.c_int
    lib.ipr_device_count.restype = C.c_int
    lib.ipr_na_real.restype = C.c_double
    lib.ipr_idw_f64.restype = C.c_int
    lib.ipr_idw_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, C.c_double,
                                ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_idw_kriging_f64.restype = C.c_int
    lib.ipr_idw_kriging_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32,
                                        ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_idw_kriging.restype = C.c_int
    lib.ipr_plan_idw_kriging.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_int64, C.c_char_p, C.c_size_t]
    lib.ipr_kriging_f64.restype = C.c_int
    lib.ipr_kriging_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, vp, vp, vp, vp,
                                    ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_kriging.restype = C.c_int
    lib.ipr_plan_kriging.argtypes = [vp, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, C.c_int64, C.c_char_p, C.c_size_t]
    lib.ipr_cressman_f64.restype = C.c_int
    lib.ipr_cressman_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32,
                                     ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_idw_post_f64.restype = C.c_int
    lib.ipr_idw_post_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, C.c_double,
                                     C.c_int32, C.c_int32, vp, ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_cressman_post_f64.restype = C.c_int
    lib.ipr_cressman_post_f64.argtypes = [vp, vp, C.c_int64, vp, vp, C.c_int32, vp, C.c_int32, vp, C.c_int32,
                                          C.c_int32, C.c_int32, vp, ip, C.c_int, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_set_post.restype = C.c_int
    lib.ipr_plan_set_post.argtypes = [vp, C.c_int32, C.c_int32, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_create.restype = C.c_int
    lib.ipr_plan_create.argtypes = [C.POINTER(vp), C.c_int, vp, vp, C.c_int64, vp, vp, C.c_int32, C.c_int32,
                                    C.c_char_p, C.c_size_t]
    lib.ipr_plan_destroy.restype = None
    lib.ipr_plan_destroy.argtypes = [vp]
    lib.ipr_plan_set_obs.restype = C.c_int
    lib.ipr_plan_set_obs.argtypes = [vp, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_set_obs_device.restype = C.c_int
    lib.ipr_plan_set_obs_device.argtypes = [vp, vp, C.c_char_p, C.c_size_t]
    lib.ipr_plan_idw.restype = C.c_int
    lib.ipr_plan_idw.argtypes = [vp, C.c_double, C.c_int32, C.c_int32, vp, C.c_int64, C.c_char_p, C.c_size_t]
    lib.ipr_plan_cressman.restype = C.c_int
    lib.ipr_plan_cressman.argtypes = [vp, vp, C.c_int32, C.c_int32, C.c_int32, vp, C.c_int64, C.c_int64,
                                      C.c_char_p, C.c_size_t]
    lib.ipr_plan_reset_cache.restype = None
    lib.ipr_plan_reset_cache.argtypes = [vp]
    lib.ipr_plan_sync.restype = C.c_int
    lib.ipr_plan_sync.argtypes = [vp]
    lib.ipr_plan_timer_start.restype = C.c_int
    lib.ipr_plan_timer_start.argtypes = [vp]
    lib.ipr_plan_timer_stop.restype = C.c_int
    lib.ipr_plan_timer_stop.argtypes = [vp, C.POINTER(C.c_float)]
    lib.ipr_plan_launch_count.restype = C.c_int64
    lib.ipr_plan_launch_count.argtypes = [vp]
    lib.ipr_plan_stage_blocks.restype = C.c_int
    lib.ipr_plan_stage_blocks.argtypes = [vp, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
    lib.ipr_plan_profile.restype = C.c_int
    lib.ipr_plan_profile.argtypes = [vp, C.c_int]
    lib.ipr_plan_profile_read.restype = C.c_int
    lib.ipr_plan_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.c_int]
    lib.ipr_plan_zero_pairs.restype = C.c_int64
    lib.ipr_plan_zero_pairs.argtypes = [vp]
    lib.ipr_plan_stream.restype = vp
    lib.ipr_plan_stream.argtypes = [vp]
    lib.ipr_release_cached_memory.restype = None
    lib.ipr_release_cached_memory.argtypes = []
    lib.ipr_host_alloc.restype = vp
    lib.ipr_host_alloc.argtypes = [C.c_size_t]
    lib.ipr_host_free.restype = None
    lib.ipr_host_free.argtypes = [vp]
    lib.ipr_job_start_idw.restype = C.c_int
    lib.ipr_job_start_idw.argtypes = [C.POINTER(vp), *lib.ipr_idw_post_f64.argtypes]
    lib.ipr_job_start_cressman.restype = C.c_int
    lib.ipr_job_start_cressman.argtypes = [C.POINTER(vp), *lib.ipr_cressman_post_f64.argtypes]
    lib.ipr_job_wait.restype = C.c_int
    lib.ipr_job_wait.argtypes = [vp, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ipr_job_cancel.restype = None
    lib.ipr_job_cancel.argtypes = [vp]
    lib.ipr_job_finish.restype = C.c_int
    lib.ipr_job_finish.argtypes = [vp, C.c_char_p, C.c_size_t]
    lib.ipr_cell_split_bounds.restype = None
    lib.ipr_cell_split_bounds.argtypes = [C.c_int64, C.c_int, C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def check(rc: int, errbuf) -> None:
    if rc != IPR_OK:
        raise NativeError(rc, errbuf.value.decode("utf-8", "replace"))


def errbuf():
    return C.create_string_buffer(ERRLEN)
