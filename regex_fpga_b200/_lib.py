"""ctypes loader for regex_fpga_b200/lib/librfb200.so (the C ABI in include/regex_fpga_b200.h)."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RFB_LIB") or os.path.join(_HERE, "lib", "librfb200.so")   # RFB_LIB: kernel variants (tools/dev)


class rfb_match(C.Structure):
    _fields_ = [("stream", C.c_uint32), ("pos", C.c_uint32), ("state", C.c_uint32)]


class rfb_nfa_info(C.Structure):
    _fields_ = [(n, C.c_uint32) for n in (
        "n_states", "n_transitions", "n_accepting", "n_entries", "image_ok", "image_bytes", "n_sticky",
        "sticky_words", "n_slots", "n_class_sets", "bucket_bits", "n_parts")]


class rfb_batch(C.Structure):
    _fields_ = [("data", C.c_void_p), ("data_bytes", C.c_uint64), ("n_streams", C.c_uint64),
                ("stride", C.c_uint64), ("offsets", C.c_void_p), ("n_steps", C.c_uint32),
                ("steps", C.c_void_p), ("stream_id_base", C.c_uint32), ("pos_base", C.c_uint32),
                ("state_in", C.c_void_p), ("state_out", C.c_void_p), ("state_cap", C.c_uint32),
                ("reserved", C.c_uint32)]


class rfb_result(C.Structure):
    _fields_ = [("counts", C.c_void_p), ("records", C.c_void_p), ("record_capacity", C.c_uint64),
                ("n_matches", C.c_uint64), ("n_records", C.c_uint64), ("n_dropped", C.c_uint64),
                ("n_symbols", C.c_uint64), ("n_rescanned", C.c_uint64), ("gpu_ms", C.c_float),
                ("n_launches", C.c_uint32)]


# every symbol include/regex_fpga_b200.h declares: name -> (restype, argtypes)
_VP, _U32P, _U8P = C.c_void_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint8)
SIGNATURES = {
    "rfb_abi_version": (C.c_int, []),
    "rfb_ctx_create": (C.c_int, [C.c_int, C.POINTER(_VP)]),
    "rfb_ctx_destroy": (None, [_VP]),
    "rfb_last_error": (C.c_char_p, [_VP]),
    "rfb_nfa_load_coe": (C.c_int, [_VP, C.c_char_p, C.c_int64, C.POINTER(_VP)]),
    "rfb_nfa_from_entries": (C.c_int, [_VP, _U32P, C.c_size_t, C.c_int64, C.POINTER(_VP)]),
    "rfb_nfa_destroy": (None, [_VP]),
    "rfb_nfa_get_info": (C.c_int, [_VP, C.POINTER(rfb_nfa_info)]),
    "rfb_nfa_get_entries": (C.c_int, [_VP, _U32P, C.c_size_t]),
    "rfb_nfa_calibrate": (C.c_int, [_VP, _VP, C.POINTER(rfb_batch)]),
    "rfb_nfa_calibration": (C.c_int, [_VP, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]),
    "rfb_trace_load_mem": (C.c_int, [C.c_char_p, C.POINTER(_U8P), C.POINTER(C.c_size_t)]),
    "rfb_trace_write_mem": (C.c_int, [C.c_char_p, _U8P, C.c_size_t]),
    "rfb_coe_parse": (C.c_int, [C.c_char_p, C.POINTER(_U32P), C.POINTER(C.c_size_t)]),
    "rfb_coe_write": (C.c_int, [C.c_char_p, _U32P, C.c_size_t, C.c_int]),
    "rfb_coe_detect_size": (C.c_int64, [_U32P, C.c_size_t]),
    "rfb_free": (None, [_VP]),
    "rfb_image_check": (C.c_int, [_U32P, C.c_size_t, C.c_int64, C.c_int, C.c_int, C.POINTER(rfb_nfa_info)]),
    "rfb_nfa_describe": (C.c_int, [_VP, C.c_char_p, C.c_size_t]),
    "rfb_nfa_save_image": (C.c_int, [_VP, C.c_char_p]),
    "rfb_nfa_load_image": (C.c_int, [_VP, C.c_char_p, C.POINTER(_VP)]),
    "rfb_image_file_build": (C.c_int, [_U32P, C.c_size_t, C.c_int64, C.c_char_p]),
    "rfb_image_file_check": (C.c_int, [C.c_char_p, C.POINTER(rfb_nfa_info)]),
    "rfb_tb_steps": (C.c_uint32, [C.c_uint32]),
    "rfb_scan": (C.c_int, [_VP, _VP, C.POINTER(rfb_batch), C.c_uint32, C.POINTER(rfb_result)]),
    "rfb_scan_device": (C.c_int, [_VP, _VP, C.POINTER(rfb_batch), C.c_uint32, _VP, C.POINTER(rfb_result)]),
    "rfb_scan_collect": (C.c_int, [_VP, C.POINTER(rfb_result)]),
    "rfb_scan_submit": (C.c_int, [_VP, _VP, C.POINTER(rfb_batch), C.c_uint32, C.POINTER(rfb_result)]),
    "rfb_scan_wait": (C.c_int, [_VP, C.POINTER(C.POINTER(rfb_result))]),
    "rfb_group_create": (C.c_int, [C.POINTER(C.c_int), C.c_int, C.POINTER(_VP)]),
    "rfb_group_destroy": (None, [_VP]),
    "rfb_group_size": (C.c_int, [_VP]),
    "rfb_group_ctx": (_VP, [_VP, C.c_int]),
    "rfb_group_last_error": (C.c_char_p, [_VP]),
    "rfb_group_nfa_load_coe": (C.c_int, [_VP, C.c_char_p, C.c_int64, C.POINTER(_VP)]),
    "rfb_group_nfa_from_entries": (C.c_int, [_VP, _U32P, C.c_size_t, C.c_int64, C.POINTER(_VP)]),
    "rfb_group_nfa_destroy": (None, [_VP]),
    "rfb_group_nfa_member": (_VP, [_VP, C.c_int]),
    "rfb_group_scan": (C.c_int, [_VP, _VP, C.POINTER(rfb_batch), C.c_uint32, C.POINTER(rfb_result)]),
    "rfb_fpga_cycles": (C.c_int, [_VP, _VP, _U8P, _U8P, C.c_uint32, C.POINTER(C.c_uint64)]),
}

_lib = None


def load():
    """Load the shared library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "or `make -C regex_fpga_b200/csrc`.  There is no CPU or pure-Python fallback.")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib
