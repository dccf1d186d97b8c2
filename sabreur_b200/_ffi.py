"""ctypes binding of include/sabreur_gpu.h (libsabreur_gpu.so).  No fallback: a missing library or a
missing GPU raises."""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SABREUR_GPU_LIB") or os.path.join(HERE, "lib", "libsabreur_gpu.so")  # (the override: experimental builds)

SBR_OK = 0
ST_INVALID_START, ST_INVALID_SEP, ST_UNEQUAL_LEN, ST_SHORT_READ = 0x01, 0x02, 0x04, 0x08
ST_TRUNCATED, ST_REC_OVERFLOW, ST_NONCANONICAL, ST_ERROR_MASK = 0x10, 0x20, 0x40, 0x1F
FMT_AUTO, FMT_FASTA, FMT_FASTQ = 0, 1, 2
CFG_NO_LUT, CFG_GENERIC, CFG_TIMING, CFG_EXACT_SCAN, CFG_NO_FUSE, CFG_OVERLAP = 0x1, 0x2, 0x4, 0x8, 0x10, 0x20
KEY_SHORT, KEY_NONCANON = 1 << 48, 1 << 49

# every symbol include/sabreur_gpu.h declares
EXPORTS = [
    "sbr_abi_version", "sbr_device_count", "sbr_create", "sbr_destroy", "sbr_last_error", "sbr_slot_buffer",
    "sbr_submit", "sbr_wait", "sbr_demux_device", "sbr_slot_device_result", "sbr_device_join", "sbr_scan_stats", "sbr_scan_geometry", "sbr_lut_entries", "sbr_lut_copy",
    "sbr_debug_regions", "sbr_timing_collect", "sbr_synth_table", "sbr_synth_record_bytes", "sbr_synth_host", "sbr_synth_device",
    "sbr_dev_alloc", "sbr_dev_free", "sbr_dev_upload", "sbr_dev_download", "sbr_host_alloc_pinned",
    "sbr_host_free_pinned",
]


class SbrConfig(C.Structure):
    _fields_ = [("device", C.c_int32), ("fmt", C.c_uint32), ("n_barcodes", C.c_uint32), ("bc_len", C.c_uint32),
                ("barcodes", C.c_void_p), ("bc_lens", C.c_void_p), ("mismatch", C.c_uint32), ("n_slots", C.c_uint32),
                ("batch_bytes", C.c_uint64), ("max_records", C.c_uint32), ("flags", C.c_uint32)]


class SbrResult(C.Structure):
    _fields_ = [("batch_seq", C.c_uint64), ("consumed", C.c_uint64), ("n_records", C.c_uint32),
                ("status_flags", C.c_uint32), ("first_bad_record", C.c_uint32), ("first_bad_flags", C.c_uint32),
                ("fmt", C.c_uint32), ("n_bins", C.c_uint32), ("rec_off", C.c_void_p), ("assign", C.c_void_p),
                ("bin_count", C.c_void_p), ("bin_start", C.c_void_p), ("order", C.c_void_p), ("rec_key", C.c_void_p),
                ("scan_path", C.c_uint32), ("reserved", C.c_uint32)]


class SbrSynthSpec(C.Structure):
    _fields_ = [("seed", C.c_uint64), ("fmt", C.c_uint32), ("read_len", C.c_uint32), ("mate", C.c_uint32),
                ("n_barcodes", C.c_uint32), ("bc_len", C.c_uint32), ("reserved", C.c_uint32), ("barcodes", C.c_void_p)]


_lib = None


def load():
    """dlopen the C-ABI library and declare its prototypes.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m sabreur_b200.build` (needs nvcc). "
            "sabreur_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    proto = {
        "sbr_abi_version": (i32, []),
        "sbr_device_count": (i32, []),
        "sbr_create": (i32, [C.POINTER(vp), C.POINTER(SbrConfig)]),
        "sbr_destroy": (None, [vp]),
        "sbr_last_error": (C.c_char_p, [vp]),
        "sbr_slot_buffer": (i32, [vp, u32, C.POINTER(vp), C.POINTER(u64)]),
        "sbr_submit": (i32, [vp, u32, u64, u64, u64, i32]),
        "sbr_wait": (i32, [vp, u32, C.POINTER(SbrResult)]),
        "sbr_demux_device": (i32, [vp, u32, vp, u64, i32, vp]),
        "sbr_slot_device_result": (i32, [vp, u32, C.POINTER(SbrResult)]),
        "sbr_device_join": (i32, [vp, vp]),
        "sbr_scan_stats": (i32, [vp, C.POINTER(u64 * 2), i32]),
        "sbr_scan_geometry": (i32, [vp, u32, C.POINTER(u32), C.POINTER(u32)]),
        "sbr_lut_entries": (u64, [vp]),
        "sbr_lut_copy": (i32, [vp, vp, u64]),
        "sbr_debug_regions": (i32, [vp, u32, vp, u32]),
        "sbr_timing_collect": (i32, [vp, C.POINTER(C.c_double * 3), C.POINTER(u64), C.POINTER(u64)]),
        "sbr_synth_table": (i32, [u64, u32, u32, u32, vp]),
        "sbr_synth_record_bytes": (u32, [C.POINTER(SbrSynthSpec)]),
        "sbr_synth_host": (i32, [C.POINTER(SbrSynthSpec), u64, u64, vp]),
        "sbr_synth_device": (i32, [i32, C.POINTER(SbrSynthSpec), u64, u64, vp, vp]),
        "sbr_dev_alloc": (i32, [i32, u64, C.POINTER(vp)]),
        "sbr_dev_free": (i32, [i32, vp]),
        "sbr_dev_upload": (i32, [i32, vp, vp, u64]),
        "sbr_dev_download": (i32, [i32, vp, vp, u64]),
        "sbr_host_alloc_pinned": (i32, [u64, C.POINTER(vp)]),
        "sbr_host_free_pinned": (i32, [vp]),
    }
    for name, (res, args) in proto.items():
        fn = getattr(L, name)  # AttributeError if the library does not export a declared symbol
        fn.restype, fn.argtypes = res, args
    _lib = L
    return L
