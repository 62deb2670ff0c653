"""ctypes binding of the C ABI declared in include/fmradion_b200.h.

The shared library is built in-tree by __graft_entry__.build() (nvcc, sm_100a). There is no
CPU fallback: importing this module without the built library, or creating a handle
without a CUDA device, raises.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfmradion_b200.so")

FMR_OK = 0
STATUS_NAMES = {0: "FMR_OK", 1: "FMR_ERR_INVALID", 2: "FMR_ERR_UNSUPPORTED", 3: "FMR_ERR_CUDA",
                4: "FMR_ERR_CAPACITY"}


class FmrError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("%s: %s" % (STATUS_NAMES.get(status, status), msg))
        self.status = status


class FmConfig(C.Structure):
    _fields_ = [("input_rate", C.c_double), ("fs4_shift", C.c_int), ("fmfilter", C.c_int),
                ("stereo", C.c_int), ("deemphasis_us", C.c_double), ("pilot_shift", C.c_int),
                ("multipath_stages", C.c_uint32), ("n_channels", C.c_uint32),
                ("max_samples_per_call", C.c_uint32), ("max_blocks_per_call", C.c_uint32),
                ("device", C.c_int), ("fmfilter_coeff", C.c_void_p), ("fmfilter_ntaps", C.c_uint32)]


class FmStats(C.Structure):
    _fields_ = [("stereo_detected", C.c_int), ("tuning_offset", C.c_float),
                ("baseband_level", C.c_float), ("pilot_level", C.c_double), ("if_rms", C.c_float),
                ("multipath_error", C.c_double), ("if_agc_gain", C.c_float), ("pll_freq", C.c_double),
                ("pll_phase", C.c_double), ("pll_lock_cnt", C.c_int), ("decoder_calls", C.c_uint64),
                ("n_pps", C.c_uint32)]


class PpsEvent(C.Structure):
    _fields_ = [("pps_index", C.c_uint64), ("sample_index", C.c_uint64),
                ("block_position", C.c_double), ("block", C.c_uint32)]


class AmConfig(C.Structure):
    _fields_ = [("input_rate", C.c_double), ("fs4_shift", C.c_int), ("amfilter", C.c_int),
                ("mode", C.c_int), ("n_channels", C.c_uint32), ("max_samples_per_call", C.c_uint32),
                ("max_blocks_per_call", C.c_uint32), ("device", C.c_int), ("amfilter_coeff", C.c_void_p),
                ("amfilter_ntaps", C.c_uint32), ("nbfm_freq_dev", C.c_double)]


class AmStats(C.Structure):
    _fields_ = [("baseband_level", C.c_double), ("af_agc_gain", C.c_float),
                ("if_agc_gain", C.c_float), ("if_rms", C.c_float), ("decoder_calls", C.c_uint64),
                ("tuning_offset", C.c_float)]


class OutputConfig(C.Structure):
    _fields_ = [("out_format", C.c_int), ("squelch_level", C.c_double), ("gain", C.c_double)]


class BlockLevel(C.Structure):
    _fields_ = [("if_rms", C.c_float), ("audio_mean", C.c_float), ("audio_rms", C.c_float), ("gain", C.c_float)]


# enum values of include/fmradion_b200.h
IQ_CF32, IQ_S16, IQ_S8, IQ_U8, IQ_S24 = range(5)
OUT_F64, OUT_F32, OUT_S16 = range(3)
IQ_BYTES = {IQ_CF32: 8, IQ_S16: 4, IQ_S8: 2, IQ_U8: 2, IQ_S24: 6}

# every symbol include/fmradion_b200.h declares
EXPORTS = [
    "fmr_last_error", "fmr_version", "fmr_device_sm_count",
    "fmr_fm_create", "fmr_fm_destroy", "fmr_fm_process_host", "fmr_fm_process_device",
    "fmr_fm_process_host_i16", "fmr_fm_process_device_i16",
    "fmr_fm_query_output", "fmr_fm_schedule", "fmr_am_schedule", "fmr_fm_stats", "fmr_fm_pps_events", "fmr_fm_coeffs",
    "fmr_fm_block_flags", "fmr_fm_tap_if", "fmr_fm_last_launches", "fmr_fm_last_plan", "fmr_fm_describe", "fmr_fm_set_profiling",
    "fmr_fm_stage_times",
    "fmr_am_create", "fmr_am_destroy", "fmr_am_process_host", "fmr_am_process_device",
    "fmr_am_query_output", "fmr_am_stats", "fmr_am_last_launches", "fmr_am_set_profiling",
    "fmr_am_stage_times",
    "fmr_fm_process_host_io", "fmr_fm_process_device_io", "fmr_fm_block_levels",
    "fmr_am_process_host_io", "fmr_am_process_device_io", "fmr_am_block_levels",
]

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            "%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp, u32p = C.c_void_p, C.POINTER(C.c_uint32)
    L.fmr_last_error.restype = C.c_char_p
    L.fmr_version.restype = C.c_char_p
    L.fmr_device_sm_count.argtypes = [C.c_int]
    L.fmr_fm_create.argtypes = [C.POINTER(FmConfig), C.POINTER(vp)]
    L.fmr_fm_destroy.argtypes = [vp]
    L.fmr_fm_destroy.restype = None
    proc_host = [vp, vp, C.c_size_t, vp, C.c_uint32, vp, C.c_size_t, vp]
    L.fmr_fm_process_host.argtypes = proc_host
    L.fmr_fm_process_device.argtypes = proc_host + [vp]
    L.fmr_fm_process_host_i16.argtypes = proc_host
    L.fmr_fm_process_device_i16.argtypes = proc_host + [vp]
    L.fmr_fm_query_output.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_uint64), vp]
    L.fmr_fm_stats.argtypes = [vp, C.c_uint32, C.POINTER(FmStats)]
    L.fmr_fm_pps_events.argtypes = [vp, C.c_uint32, vp, C.c_uint32, u32p]
    L.fmr_fm_coeffs.argtypes = [vp, C.c_uint32, vp, C.c_size_t]
    L.fmr_fm_block_flags.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    L.fmr_fm_tap_if.argtypes = [vp, C.c_uint32, vp, C.c_size_t, C.POINTER(C.c_uint64)]
    L.fmr_fm_last_launches.argtypes = [vp]
    L.fmr_fm_last_launches.restype = C.c_uint32
    L.fmr_fm_last_plan.argtypes = [vp, C.POINTER(C.c_uint64)]
    L.fmr_fm_describe.argtypes = [vp, C.c_char_p, C.c_size_t]
    L.fmr_fm_describe.restype = C.c_size_t
    L.fmr_fm_set_profiling.argtypes = [vp, C.c_int]
    L.fmr_fm_stage_times.argtypes = [vp, vp, vp, C.c_uint32, u32p]
    L.fmr_am_set_profiling.argtypes = [vp, C.c_int]
    L.fmr_am_stage_times.argtypes = [vp, vp, vp, C.c_uint32, u32p]
    L.fmr_fm_schedule.argtypes = [C.c_double, C.c_int, C.c_uint64, vp, C.c_uint32, vp, vp]
    L.fmr_am_schedule.argtypes = [C.c_double, C.c_uint64, vp, C.c_uint32, vp]
    L.fmr_am_create.argtypes = [C.POINTER(AmConfig), C.POINTER(vp)]
    L.fmr_am_destroy.argtypes = [vp]
    L.fmr_am_destroy.restype = None
    L.fmr_am_process_host.argtypes = proc_host
    L.fmr_am_process_device.argtypes = proc_host + [vp]
    L.fmr_am_query_output.argtypes = [vp, vp, C.c_uint32, C.POINTER(C.c_uint64), vp]
    L.fmr_am_stats.argtypes = [vp, C.c_uint32, C.POINTER(AmStats)]
    L.fmr_am_last_launches.argtypes = [vp]
    L.fmr_am_last_launches.restype = C.c_uint32
    proc_io = [vp, vp, C.c_int, C.c_size_t, vp, C.c_uint32, C.POINTER(OutputConfig), vp, C.c_size_t, vp]
    L.fmr_fm_process_host_io.argtypes = proc_io
    L.fmr_fm_process_device_io.argtypes = proc_io + [vp]
    L.fmr_fm_block_levels.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    L.fmr_am_process_host_io.argtypes = proc_io
    L.fmr_am_process_device_io.argtypes = proc_io + [vp]
    L.fmr_am_block_levels.argtypes = [vp, C.c_uint32, vp, C.c_uint32]
    _lib = L
    return L


def check(status):
    if status != FMR_OK:
        raise FmrError(status, lib().fmr_last_error().decode())
