"""ctypes binding of the CPU oracle -- TEST INFRASTRUCTURE ONLY.

Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs.  The product package (regex_fpga_b200) never imports this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

REC_DTYPE = np.dtype([("stream", "<u4"), ("pos", "<u4"), ("state", "<u4")])


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        u32p, u8p, u64p, u16p = (C.POINTER(C.c_uint32), C.POINTER(C.c_uint8), C.POINTER(C.c_uint64),
                                 C.POINTER(C.c_uint16))
        L.orc_coe_parse.argtypes = [C.c_char_p, C.POINTER(u32p), C.POINTER(C.c_size_t)]
        L.orc_detect_size.argtypes = [u32p, C.c_size_t]
        L.orc_detect_size.restype = C.c_int64
        L.orc_mem_parse.argtypes = [C.c_char_p, C.POINTER(u8p), C.POINTER(C.c_size_t)]
        L.orc_free.argtypes = [C.c_void_p]
        L.orc_b_scan.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, C.c_uint64, C.c_uint32, u64p,
                                 C.c_void_p, C.c_uint64, u64p, u64p, u32p]
        L.orc_b_scan_many.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, C.c_uint64, C.c_uint64,
                                      C.c_uint64, C.c_int, u64p, C.c_void_p, C.c_uint64, u64p, u64p]
        L.orc_a_run.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, u8p, C.c_uint64, C.c_int, C.c_int,
                                u16p, u16p, u64p, u64p, C.c_void_p, C.c_uint64, u64p, u64p]
        L.orc_cycle_model.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, u8p, C.c_uint64, u64p]
        L.orc_a_run_many.argtypes = [u32p, C.c_size_t, C.c_uint32, u8p, C.c_uint64, C.c_uint64,
                                     C.c_uint64, C.c_int, C.c_int, u64p, u64p, u64p]
        _LIB = L
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def coe_parse(path):
    L = lib()
    ptr = C.POINTER(C.c_uint32)()
    n = C.c_size_t()
    rc = L.orc_coe_parse(path.encode(), C.byref(ptr), C.byref(n))
    if rc:
        raise ValueError(f"orc_coe_parse({path}) -> {rc}")
    out = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()
    L.orc_free(ptr)
    return out


def detect_size(E):
    E = np.ascontiguousarray(E, dtype=np.uint32)
    return int(lib().orc_detect_size(_p(E, C.c_uint32), E.size))


def mem_parse(path):
    L = lib()
    ptr = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    rc = L.orc_mem_parse(path.encode(), C.byref(ptr), C.byref(n))
    if rc:
        raise ValueError(f"orc_mem_parse({path}) -> {rc}")
    out = np.ctypeslib.as_array(ptr, shape=(n.value,)).copy()
    L.orc_free(ptr)
    return out


def b_scan(E, size, data, n_steps, stream_id=0, cap=1 << 20):
    """Functional oracle, one stream.  Returns dict(counts, recs, n_recs, sum_active, max_active)."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    data = np.ascontiguousarray(data, dtype=np.uint8)
    assert data.size >= n_steps
    counts = np.zeros(size, dtype=np.uint64)
    recs = np.zeros(cap, dtype=REC_DTYPE)
    nr, sa, ma = C.c_uint64(), C.c_uint64(), C.c_uint32()
    rc = lib().orc_b_scan(_p(E, C.c_uint32), E.size, size, _p(data, C.c_uint8), n_steps, stream_id,
                          _p(counts, C.c_uint64), recs.ctypes.data, cap, C.byref(nr), C.byref(sa),
                          C.byref(ma))
    if rc:
        raise ValueError(f"orc_b_scan -> {rc}")
    return dict(counts=counts, recs=recs[: min(nr.value, cap)], n_recs=nr.value,
                sum_active=sa.value, max_active=ma.value)


def b_scan_many(E, size, data, n_streams, stride, n_steps, n_threads=0, cap=1 << 22, want_recs=True):
    """Functional oracle over n_streams streams at data[s*stride:], canonical record order."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    assert n_streams == 0 or data.size >= (n_streams - 1) * stride + n_steps
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    counts = np.zeros(size, dtype=np.uint64)
    recs = np.zeros(cap if want_recs else 0, dtype=REC_DTYPE)
    nr, sa = C.c_uint64(), C.c_uint64()
    rc = lib().orc_b_scan_many(_p(E, C.c_uint32), E.size, size, _p(data, C.c_uint8), n_streams, stride,
                               n_steps, n_threads, _p(counts, C.c_uint64),
                               recs.ctypes.data if want_recs else None, cap, C.byref(nr), C.byref(sa))
    if rc:
        raise ValueError(f"orc_b_scan_many -> {rc}")
    return dict(counts=counts, recs=recs[: min(nr.value, recs.size)], n_recs=nr.value,
                sum_active=sa.value)


def a_run(E, size, lo, hi, M, addr_bits=16, fast_idle=False, cap=1 << 20):
    """Cycle-level oracle (FPGA.v + ROM + testbench) on an M-entry trace pair."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    lo = np.ascontiguousarray(lo, dtype=np.uint8)
    hi = np.ascontiguousarray(hi, dtype=np.uint8)
    assert lo.size >= M and hi.size >= M
    mc1 = np.zeros(size, dtype=np.uint16)
    mc2 = np.zeros(size, dtype=np.uint16)
    c1 = np.zeros(size, dtype=np.uint64)
    c2 = np.zeros(size, dtype=np.uint64)
    recs = np.zeros(cap, dtype=REC_DTYPE)
    nr, cyc = C.c_uint64(), C.c_uint64()
    rc = lib().orc_a_run(_p(E, C.c_uint32), E.size, size, _p(lo, C.c_uint8), _p(hi, C.c_uint8), M,
                         addr_bits, int(fast_idle), _p(mc1, C.c_uint16), _p(mc2, C.c_uint16),
                         _p(c1, C.c_uint64), _p(c2, C.c_uint64), recs.ctypes.data, cap, C.byref(nr),
                         C.byref(cyc))
    if rc:
        raise ValueError(f"orc_a_run -> {rc}")
    return dict(mc1=mc1, mc2=mc2, counts1=c1, counts2=c2, recs=recs[: min(nr.value, cap)],
                n_recs=nr.value, cycles=cyc.value)


def cycle_model(E, size, lo, hi, M):
    E = np.ascontiguousarray(E, dtype=np.uint32)
    lo = np.ascontiguousarray(lo, dtype=np.uint8)
    hi = np.ascontiguousarray(hi, dtype=np.uint8)
    cyc = C.c_uint64()
    rc = lib().orc_cycle_model(_p(E, C.c_uint32), E.size, size, _p(lo, C.c_uint8), _p(hi, C.c_uint8), M,
                               C.byref(cyc))
    if rc:
        raise ValueError(f"orc_cycle_model -> {rc}")
    return cyc.value


def a_run_many(E, size, data, n_pairs, stride, M, n_threads=0, fast_idle=False):
    """CPU baseline: n_pairs (lo,hi) pairs through the cycle-level oracle on n_threads threads."""
    E = np.ascontiguousarray(E, dtype=np.uint32)
    data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
    assert data.size >= (2 * n_pairs - 1) * stride + M
    if n_threads <= 0:
        n_threads = os.cpu_count() or 1
    counts = np.zeros(size, dtype=np.uint64)
    cyc, sym = C.c_uint64(), C.c_uint64()
    rc = lib().orc_a_run_many(_p(E, C.c_uint32), E.size, size, _p(data, C.c_uint8), n_pairs, stride, M,
                              n_threads, int(fast_idle), _p(counts, C.c_uint64), C.byref(cyc),
                              C.byref(sym))
    if rc:
        raise ValueError(f"orc_a_run_many -> {rc}")
    return dict(counts=counts, cycles=cyc.value, symbols=sym.value, threads=n_threads)
