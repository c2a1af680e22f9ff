"""ctypes binding of oracle/_ref/libref.so -- TEST INFRASTRUCTURE ONLY.

libref.so is the reference's OWN Verilog (Design/FPGA.v, read where it lies under /root/reference, never copied)
translated to C by oracle/vsim/v2c.py and clocked by oracle/vsim/tb_driver.c exactly as the reference's testbench
does.  It exists to pin the hand-written oracles and to time "the reference on host cores"; only tests/,
bench.py's CPU legs and tests/golden/make_golden.py import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("RFB_REFERENCE", "/root/reference")
LIB_PATH = os.path.join(_HERE, "_ref", "libref.so")
_LIB = None

REC_DTYPE = np.dtype([("stream", "<u4"), ("pos", "<u4"), ("state", "<u4")])


def reference_present():
    return os.path.exists(os.path.join(REFERENCE, "Design", "FPGA.v"))


def build():
    """`make -C oracle _ref`: only where the reference sources exist (the build container)."""
    if not reference_present():
        raise RuntimeError(f"{REFERENCE}/Design/FPGA.v is missing: libref.so can only be built where the reference lies")
    subprocess.check_call(["make", "-s", "-C", _HERE, "_ref", f"REF={REFERENCE}"])


def available():
    return os.path.exists(LIB_PATH) or reference_present()


def lib():
    global _LIB
    if _LIB is None:
        if reference_present():
            build()                      # make: a no-op when up to date
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing and cannot be built here")
        L = C.CDLL(LIB_PATH)
        u32p, u8p, u64p, u16p = (C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_uint16))
        L.ref_tb_run.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, u8p, C.c_uint64, C.c_int, u16p, u16p, u64p, u64p,
                                 C.c_void_p, C.c_uint64, u64p, u64p, u64p, C.c_uint64]
        L.ref_tb_run_many.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_int,
                                      u64p, u64p, u64p]
        L.ref_tb_module.restype = C.c_char_p
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def tb_run(E, size, lo, hi, M, xfill=0, cap=1 << 20, trace_cap=0):
    """The reference testbench on an M-entry (lo, hi) trace pair (the TB hard-codes M = 200000).
    xfill: the bit every x/z (power-up register values, z literals) is replaced by in the 2-state model.
    Returns dict(mc1, mc2, counts1, counts2, recs, n_recs, cycles[, trace])."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    lo = np.ascontiguousarray(lo, dtype=np.uint8)
    hi = np.ascontiguousarray(hi, dtype=np.uint8)
    assert lo.size >= M and hi.size >= M
    mc1, mc2 = np.zeros(size, np.uint16), np.zeros(size, np.uint16)
    c1, c2 = np.zeros(size, np.uint64), np.zeros(size, np.uint64)
    recs = np.zeros(cap, dtype=REC_DTYPE)
    trace = np.zeros(max(trace_cap, 1), dtype=np.uint64)
    nr, cyc = C.c_uint64(), C.c_uint64()
    rc = lib().ref_tb_run(_p(E, C.c_uint32), E.size, size, _p(lo, C.c_uint8), _p(hi, C.c_uint8), M, xfill,
                          _p(mc1, C.c_uint16), _p(mc2, C.c_uint16), _p(c1, C.c_uint64), _p(c2, C.c_uint64),
                          recs.ctypes.data, cap, C.byref(nr), C.byref(cyc), _p(trace, C.c_uint64) if trace_cap else None, trace_cap)
    if rc:
        raise ValueError(f"ref_tb_run -> {rc}")
    out = dict(mc1=mc1, mc2=mc2, counts1=c1, counts2=c2, recs=recs[: min(nr.value, cap)], n_recs=nr.value, cycles=cyc.value)
    if trace_cap:
        out["trace"] = trace[: min(cyc.value, trace_cap)]
    return out


def tb_run_many(E, size, data, n_pairs, stride, M, n_threads=0):
    """CPU arm of bench.py: n_pairs (lo, hi) pairs through the reference design on n_threads host threads."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    assert data.size >= (2 * n_pairs - 1) * stride + M
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    counts = np.zeros(size, dtype=np.uint64)
    cyc, sym = C.c_uint64(), C.c_uint64()
    rc = lib().ref_tb_run_many(_p(E, C.c_uint32), E.size, size, _p(data, C.c_uint8), n_pairs, stride, M, n_threads,
                               _p(counts, C.c_uint64), C.byref(cyc), C.byref(sym))
    if rc:
        raise ValueError(f"ref_tb_run_many -> {rc}")
    return dict(counts=counts, cycles=cyc.value, symbols=sym.value, threads=n_threads)
