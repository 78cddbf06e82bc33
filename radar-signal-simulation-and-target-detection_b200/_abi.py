"""ctypes binding of include/rsp.h (librsp.so).  This is the same boundary a MEX gateway binds
(see INTEGRATION.md); nothing here computes -- it only marshals pointers and sizes.

The product path fails loudly when the CUDA library is missing: there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# RSP_LIBRARY selects another build of the same ABI (A/B measurements of kernel generations)
LIB_PATH = os.environ.get("RSP_LIBRARY") or os.path.join(HERE, "lib", "librsp.so")

RSP_ABI_VERSION = 1
RSP_OK = 0
RSP_ERR_INVALID_ARG, RSP_ERR_UNSUPPORTED, RSP_ERR_CUDA, RSP_ERR_NO_DEVICE, RSP_ERR_OVERFLOW, RSP_ERR_NOT_READY = \
    -1, -2, -3, -4, -5, -6
RSP_LAYOUT_PCN, RSP_LAYOUT_MATLAB = 0, 1
RSP_MEM_HOST, RSP_MEM_DEVICE = 0, 1
RSP_C64, RSP_C128 = 0, 1


class rsp_params(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("n_channels", C.c_int32), ("n_beams", C.c_int32),
                ("n_pulses", C.c_int32), ("n_samples", C.c_int32), ("seg_start", C.c_int32 * 3),
                ("n_gates", C.c_int32 * 3), ("fir_delay", C.c_int32), ("t_cfar", C.c_float),
                ("guard_r", C.c_int32), ("guard_v", C.c_int32), ("ref_r", C.c_int32), ("ref_v", C.c_int32),
                ("max_detections", C.c_int32), ("monopulse_complex", C.c_int32), ("device", C.c_int32)]


class rsp_constants(C.Structure):
    _fields_ = [("dbf_weights", C.c_void_p), ("fir", C.c_void_p), ("n_fir", C.c_int32),
                ("mf_medium", C.c_void_p), ("n_mf_medium", C.c_int32),
                ("mf_long", C.c_void_p), ("n_mf_long", C.c_int32),
                ("mtd_win", C.c_void_p), ("range_axis", C.c_void_p), ("velocity_axis", C.c_void_p),
                ("delta_r", C.c_double), ("delta_v", C.c_double),
                ("beam_angles_deg", C.c_void_p), ("k_slopes", C.c_void_p)]


class rsp_detection(C.Structure):
    _fields_ = [("v_idx", C.c_int32), ("r_idx", C.c_int32), ("pair_idx", C.c_int32), ("power", C.c_float),
                ("range", C.c_double), ("velocity", C.c_double), ("angle", C.c_double)]


class rsp_target(C.Structure):
    _fields_ = [("range", C.c_double), ("velocity", C.c_double), ("angle", C.c_double), ("power", C.c_double)]


class rsp_cluster_params(C.Structure):
    _fields_ = [("max_range_sep", C.c_double), ("max_vel_sep", C.c_double), ("max_angle_sep", C.c_double)]


class rsp_info(C.Structure):
    _fields_ = [("n_gates_total", C.c_int32), ("fft_len_medium", C.c_int32), ("fft_len_long", C.c_int32),
                ("blocks_medium", C.c_int32), ("blocks_long", C.c_int32), ("kernels_per_cpi", C.c_int32),
                ("algorithmic_bytes_per_cpi", C.c_int64), ("launches_total", C.c_int64), ("lanes", C.c_int32),
                ("graph_launches", C.c_int32)]


class rsp_target_in(C.Structure):
    _fields_ = [("range", C.c_double), ("velocity", C.c_double), ("elevation_deg", C.c_double), ("snr_db", C.c_double)]


class rsp_waveform(C.Structure):
    _fields_ = [("tx_pulse", C.c_void_p), ("c", C.c_double), ("fs", C.c_double), ("wavelength", C.c_double),
                ("prt", C.c_double), ("element_spacing", C.c_double), ("p_signal_unscaled", C.c_double)]


class rsp_stage2_config(C.Structure):
    _fields_ = [("pulse", C.c_void_p * 3), ("n_pulse", C.c_int32 * 3), ("mtd_win", C.c_void_p),
                ("zero_vel_bins", C.c_int32)]


class rsp_cfar1d_params(C.Structure):
    _fields_ = [("ref_cells", C.c_int32), ("save_cells", C.c_int32), ("method", C.c_int32), ("zero_vel_bins", C.c_int32),
                ("t_cfar", C.c_float), ("seg_len", C.c_int32 * 3)]


class rsp_kernel_times(C.Structure):
    _fields_ = [("n", C.c_int32), ("name", C.c_char_p * 12), ("total_ms", C.c_double * 12), ("launches", C.c_int64 * 12)]


# numpy dtype with the exact memory layout of rsp_detection (40 bytes)
DETECTION_DTYPE = [("v_idx", "<i4"), ("r_idx", "<i4"), ("pair_idx", "<i4"), ("power", "<f4"),
                   ("range", "<f8"), ("velocity", "<f8"), ("angle", "<f8")]
TARGET_DTYPE = [("range", "<f8"), ("velocity", "<f8"), ("angle", "<f8"), ("power", "<f8")]

# every symbol include/rsp.h declares: (name, restype, argtypes)
_P = C.c_void_p
SYMBOLS = [
    ("rsp_abi_version", C.c_int, []),
    ("rsp_device_count", C.c_int, []),
    ("rsp_create", C.c_int, [C.POINTER(rsp_params), C.POINTER(_P)]),
    ("rsp_destroy", None, [_P]),
    ("rsp_last_error", C.c_char_p, [_P]),
    ("rsp_upload_constants", C.c_int, [_P, C.POINTER(rsp_constants)]),
    ("rsp_set_stream", C.c_int, [_P, _P]),
    ("rsp_synchronize", C.c_int, [_P]),
    ("rsp_process_cpi", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, C.c_int, _P, C.c_int32, C.POINTER(C.c_int32)]),
    ("rsp_stream_enqueue", C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32]),
    ("rsp_submit_cpi", C.c_int, [_P, _P, _P, C.c_int32]),
    ("rsp_stream_slots", C.c_int, [_P]),
    ("rsp_stream_device_buffers", C.c_int, [_P, C.POINTER(_P), C.POINTER(_P)]),
    ("rsp_stream_fetch", C.c_int, [_P, C.c_int32, _P, C.c_int32, C.POINTER(C.c_int32)]),
    ("rsp_sort_detections", C.c_int, [_P, C.c_int32]),
    ("rsp_get_beam", C.c_int, [_P, _P]),
    ("rsp_get_pc", C.c_int, [_P, _P]),
    ("rsp_get_rdm", C.c_int, [_P, _P]),
    ("rsp_get_amp", C.c_int, [_P, _P]),
    ("rsp_cluster", C.c_int, [_P, C.c_int32, C.POINTER(rsp_cluster_params), _P, C.POINTER(C.c_int32), _P,
                              C.POINTER(C.c_int32)]),
    ("rsp_process_frame", C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.POINTER(rsp_cluster_params), _P, C.c_int32,
                                    C.POINTER(C.c_int32)]),
    ("rsp_stage2_configure", C.c_int, [_P, C.POINTER(rsp_stage2_config)]),
    ("rsp_stage2_mtd", C.c_int, [_P, _P, C.c_int, _P, _P]),
    ("rsp_cfar1d", C.c_int, [C.c_int32, _P, C.c_int32, C.c_int32, C.c_int32, C.POINTER(rsp_cfar1d_params), _P, _P]),
    ("rsp_stage2_cfar", C.c_int, [_P, C.POINTER(rsp_cfar1d_params), _P, _P]),
    ("rsp_set_waveform", C.c_int, [_P, C.POINTER(rsp_waveform)]),
    ("rsp_synthesize", C.c_int, [_P, _P, C.c_int32, C.c_double, C.c_uint64, _P]),
    ("rsp_process_targets", C.c_int, [_P, _P, C.c_int32, C.c_double, C.c_uint64, C.POINTER(rsp_cluster_params), _P, C.c_int32,
                                      C.POINTER(C.c_int32), _P, C.c_int32, C.POINTER(C.c_int32)]),
    ("rsp_submit_targets", C.c_int, [_P, _P, C.c_int32, C.c_double, C.c_uint64, C.c_int32]),
    ("rsp_process_frames", C.c_int, [_P, _P, _P, C.c_int32, C.c_double, _P, C.POINTER(rsp_cluster_params), C.c_int32, C.c_int32,
                                     _P, C.c_int32, _P, _P, C.c_int64, _P]),
    ("rsp_fetch_targets", C.c_int, [_P, C.c_int32, C.POINTER(rsp_cluster_params), _P, C.c_int32, C.POINTER(C.c_int32), _P, C.c_int32,
                                    C.POINTER(C.c_int32)]),
    ("rsp_get_info", C.c_int, [_P, C.POINTER(rsp_info)]),
    ("rsp_set_profiling", C.c_int, [_P, C.c_int]),
    ("rsp_get_kernel_times", C.c_int, [_P, C.POINTER(rsp_kernel_times)]),
    ("rsp_get_fused_trace", C.c_int, [_P, _P, C.c_int32, C.POINTER(C.c_int32)]),
]

_lib = None


class RspError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"librsp error {code}: {msg}")
        self.code = code


def load() -> C.CDLL:
    """dlopen librsp.so and type every entry point.  Raises if the library has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` (nvcc, sm_100a). "
            "The CUDA library is the product path; there is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    for name, restype, argtypes in SYMBOLS:
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(rc: int, ctx=None) -> None:
    if rc != RSP_OK:
        msg = load().rsp_last_error(ctx)
        raise RspError(rc, msg.decode("utf-8", "replace") if msg else "")
