"""regex_fpga_b200 -- B200-native CSR-NFA scan engine (drop-in for the Regex-FPGA hot path).

Everything that computes lives in csrc/ (CUDA sm_100a + C++ host) behind the C ABI declared in
include/regex_fpga_b200.h; this package is the thin ctypes mirror of that ABI.
"""
from .engine import (Context, Nfa, Group, GroupNfa, RfbError, ScanResult, MATCH_DTYPE, SCAN_SORT_RECORDS, SCAN_FORCE_WARP,
                     SCAN_NO_COUNTS, SCAN_ASYNC, SCAN_ACCUMULATE, STATE_OVERFLOW, coe_parse, coe_write, coe_detect_size,
                     trace_load_mem, trace_write_mem, tb_steps, image_check, image_file_build,
                     image_file_check)

__all__ = ["Context", "Nfa", "Group", "GroupNfa", "RfbError", "ScanResult", "MATCH_DTYPE", "SCAN_SORT_RECORDS", "SCAN_FORCE_WARP",
           "SCAN_NO_COUNTS", "SCAN_ASYNC", "SCAN_ACCUMULATE", "STATE_OVERFLOW", "coe_parse", "coe_write", "coe_detect_size",
           "trace_load_mem", "trace_write_mem", "tb_steps", "image_check", "image_file_build", "image_file_check"]
