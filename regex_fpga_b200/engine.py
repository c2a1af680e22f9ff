"""Host-side mirror of the C ABI: Context / Nfa / scan().

The reference's only "API" is the port list of `top` (Design/top.v:1-3) driven by the testbench
(Simulation/testbench_BLK_Mem.sv): load a BRAM image, feed two byte traces, count accept pulses per
state.  These classes expose exactly that through librfb200.so; all compute happens in the CUDA
kernels -- there is no Python or CPU implementation of the scan anywhere in this package.
"""
import ctypes as C
import weakref

import numpy as np

from . import _lib
from ._lib import rfb_batch, rfb_nfa_info, rfb_result

MATCH_DTYPE = np.dtype([("stream", "<u4"), ("pos", "<u4"), ("state", "<u4")])

SCAN_SORT_RECORDS = 1
SCAN_FORCE_WARP = 2
SCAN_NO_COUNTS = 4
SCAN_ASYNC = 8
SCAN_ACCUMULATE = 16


class RfbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"librfb200 error {code}: {msg}")
        self.code = code


def _check(rc, ctx=None):
    if rc != 0:
        msg = _lib.load().rfb_last_error(ctx).decode(errors="replace")
        raise RfbError(rc, msg)


def _info_dict(info):
    return {n: int(getattr(info, n)) for n, _ in rfb_nfa_info._fields_}


# ---- host-only helpers (formats; no GPU) ---------------------------------------------------------
def coe_parse(path):
    """Block_Mem/*.coe -> uint32 entries[4*line+slot] (slot 0 = rd_bus[127:96], Design/FPGA.v:884)."""
    L = _lib.load()
    p = C.POINTER(C.c_uint32)()
    n = C.c_size_t()
    _check(L.rfb_coe_parse(str(path).encode(), C.byref(p), C.byref(n)))
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy()
    L.rfb_free(p)
    return out


def coe_write(path, entries, style=0):
    e = np.ascontiguousarray(entries, dtype=np.uint32)
    _check(_lib.load().rfb_coe_write(str(path).encode(), e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size, style))


def coe_detect_size(entries):
    e = np.ascontiguousarray(entries, dtype=np.uint32)
    return int(_lib.load().rfb_coe_detect_size(e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size))


def trace_load_mem(path):
    """Simulation/*.mem ($readmemh text) -> uint8 array (testbench_BLK_Mem.sv:34-35)."""
    L = _lib.load()
    p = C.POINTER(C.c_uint8)()
    n = C.c_size_t()
    _check(L.rfb_trace_load_mem(str(path).encode(), C.byref(p), C.byref(n)))
    out = np.ctypeslib.as_array(p, shape=(n.value,)).copy() if n.value else np.zeros(0, np.uint8)
    L.rfb_free(p)
    return out


def trace_write_mem(path, data):
    d = np.ascontiguousarray(data, dtype=np.uint8)
    _check(_lib.load().rfb_trace_write_mem(str(path).encode(), d.ctypes.data_as(C.POINTER(C.c_uint8)), d.size))


def tb_steps(trace_entries):
    """Symbol steps the testbench executes on an M-entry trace: M-1 (testbench_BLK_Mem.sv:71-86)."""
    return int(_lib.load().rfb_tb_steps(trace_entries))


def image_check(entries, n_states=-1, sticky_words=0, bucket_bits=0):
    """Host-only: build the execution image and verify it against the CSR for all (state, symbol)."""
    e = np.ascontiguousarray(entries, dtype=np.uint32)
    info = rfb_nfa_info()
    _check(_lib.load().rfb_image_check(e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size, n_states, sticky_words,
                                       bucket_bits, C.byref(info)))
    return _info_dict(info)


def image_file_build(entries, path, n_states=-1):
    """Host-only: build the scan plan (parts + execution images) of an NFA and write it as an image file."""
    e = np.ascontiguousarray(entries, dtype=np.uint32)
    _check(_lib.load().rfb_image_file_build(e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size, n_states, str(path).encode()))


def image_file_check(path):
    """Host-only: read an image file and verify every table against the CSR it carries; returns the info dict."""
    info = rfb_nfa_info()
    _check(_lib.load().rfb_image_file_check(str(path).encode(), C.byref(info)))
    return _info_dict(info)


# ---- GPU objects -----------------------------------------------------------------------------------
class Context:
    """One GPU (one process per GPU).  Fails loudly when no CUDA device is usable."""

    def __init__(self, device_id=0):
        self._L = _lib.load()
        h = C.c_void_p()
        _check(self._L.rfb_ctx_create(device_id, C.byref(h)))
        self._h = h
        self.device_id = device_id
        self._nfas = weakref.WeakSet()
        self._inflight = []          # (result, arrays...) of submitted batches, oldest first

    def close(self):
        if getattr(self, "_h", None):
            for n in list(self._nfas):      # NFAs hold device memory of this context: release them first
                n.close()
            self._L.rfb_ctx_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def load_coe(self, path, n_states=-1):
        h = C.c_void_p()
        _check(self._L.rfb_nfa_load_coe(self._h, str(path).encode(), n_states, C.byref(h)), self._h)
        return Nfa(self, h)

    def load_image(self, path):
        """An NFA from an execution-image file (Nfa.save_image / image_file_build): verified, not rebuilt."""
        h = C.c_void_p()
        _check(self._L.rfb_nfa_load_image(self._h, str(path).encode(), C.byref(h)), self._h)
        return Nfa(self, h)

    def nfa_from_entries(self, entries, n_states=-1):
        e = np.ascontiguousarray(entries, dtype=np.uint32)
        h = C.c_void_p()
        _check(self._L.rfb_nfa_from_entries(self._h, e.ctypes.data_as(C.POINTER(C.c_uint32)), e.size, n_states,
                                            C.byref(h)), self._h)
        return Nfa(self, h)


STATE_OVERFLOW = 0xFFFFFFFF


class ScanResult:
    def __init__(self, counts, records, res, state=None):
        self.counts = counts
        self.records = records
        self.state = state      # (n_streams, 1 + state_cap) uint32: [count, ids...] per stream, when requested
        self.n_matches = int(res.n_matches)
        self.n_records = int(res.n_records)
        self.n_dropped = int(res.n_dropped)
        self.n_symbols = int(res.n_symbols)
        self.n_rescanned = int(res.n_rescanned)
        self.gpu_ms = float(res.gpu_ms)
        self.n_launches = int(res.n_launches)


class Nfa:
    """A CSR NFA resident on the context's GPU (BRAM image + execution image)."""

    def __init__(self, ctx, handle):
        self.ctx = ctx
        self._L = ctx._L
        self._h = handle
        ctx._nfas.add(self)
        info = rfb_nfa_info()
        _check(self._L.rfb_nfa_get_info(self._h, C.byref(info)))
        self.info = _info_dict(info)
        self.n_states = self.info["n_states"]

    def close(self):
        if getattr(self, "_h", None):
            self._L.rfb_nfa_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def entries(self):
        out = np.zeros(self.info["n_entries"], dtype=np.uint32)
        _check(self._L.rfb_nfa_get_entries(self._h, out.ctypes.data_as(C.POINTER(C.c_uint32)), out.size))
        return out

    # -- host buffers in, host results out (H2D + kernels + D2H inside the call) --
    def describe(self):
        """One line per part: kernel, table sizes, sticky states, start-DFA size (rfb_nfa_describe)."""
        buf = C.create_string_buffer(1 << 16)
        _check(self._L.rfb_nfa_describe(self._h, buf, len(buf)), self.ctx._h)
        return buf.value.decode()

    def save_image(self, path):
        _check(self._L.rfb_nfa_save_image(self._h, str(path).encode()), self.ctx._h)

    def scan(self, data, n_streams, n_steps=0, stride=0, offsets=None, steps=None, record_capacity=1 << 20,
             flags=SCAN_SORT_RECORDS, stream_id_base=0, want_counts=True, state_in=None, want_state=False,
             state_cap=63, pos_base=0, records_out=None, counts_out=None):
        """records_out / counts_out: caller-owned result arrays (e.g. views of pinned memory, reused across calls)
        instead of fresh numpy arrays; records_out fixes record_capacity to its length.
        state_in / want_state: resumable scans -- pass the `state` of the previous call's result to continue the
        same streams (and pos_base = symbols already consumed) instead of starting from the reset state {0}."""
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        b = rfb_batch()
        b.data = data.ctypes.data if data.size else None
        b.data_bytes = data.size
        b.n_streams = n_streams
        b.stride = stride
        b.n_steps = n_steps
        b.stream_id_base = stream_id_base
        keep = [data]
        if offsets is not None:
            offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
            assert offsets.size == n_streams
            b.offsets = offsets.ctypes.data
            keep.append(offsets)
        if steps is not None:
            steps = np.ascontiguousarray(steps, dtype=np.uint32)
            assert steps.size == n_streams
            b.steps = steps.ctypes.data
            keep.append(steps)
        b.pos_base = pos_base
        state_out = None
        if state_in is not None:
            state_in = np.ascontiguousarray(state_in, dtype=np.uint32)
            assert state_in.shape[0] == n_streams and state_in.ndim == 2
            state_cap = state_in.shape[1] - 1
            b.state_in = state_in.ctypes.data
            keep.append(state_in)
        if want_state:
            state_out = np.zeros((n_streams, 1 + state_cap), dtype=np.uint32)
            b.state_out = state_out.ctypes.data
        if state_in is not None or want_state:
            b.state_cap = state_cap
        if counts_out is not None:
            assert counts_out.dtype == np.uint64 and counts_out.size >= self.n_states and counts_out.flags.c_contiguous
            want_counts = True
        counts = counts_out if counts_out is not None else (np.zeros(self.n_states, dtype=np.uint64) if want_counts else None)
        if records_out is not None:
            assert records_out.dtype == MATCH_DTYPE and records_out.flags.c_contiguous
            record_capacity = records_out.size
        records = records_out if records_out is not None else np.empty(record_capacity, dtype=MATCH_DTYPE)
        r = rfb_result()
        r.counts = counts.ctypes.data if want_counts else None
        r.records = records.ctypes.data if record_capacity else None
        r.record_capacity = record_capacity
        _check(self._L.rfb_scan(self.ctx._h, self._h, C.byref(b), flags, C.byref(r)), self.ctx._h)
        return ScanResult(counts, records[: r.n_records], r, state_out)

    # -- pipelined host path: two batches in flight (rfb_scan_submit / rfb_scan_wait) --
    def submit(self, data, n_streams, n_steps, stride, records_out, counts_out=None, flags=SCAN_SORT_RECORDS,
               stream_id_base=0, pos_base=0):
        """Enqueue the copy and the scan of a uniformly strided host batch and return at once; `wait()` completes
        the oldest submitted batch.  data / records_out / counts_out must stay alive (and should be pinned) until
        then; at most two batches may be in flight."""
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        assert records_out.dtype == MATCH_DTYPE and records_out.flags.c_contiguous
        b = rfb_batch()
        b.data = data.ctypes.data if data.size else None
        b.data_bytes = data.size
        b.n_streams = n_streams
        b.stride = stride
        b.n_steps = n_steps
        b.stream_id_base = stream_id_base
        b.pos_base = pos_base
        r = rfb_result()
        if counts_out is not None:
            assert counts_out.dtype == np.uint64 and counts_out.size >= self.n_states and counts_out.flags.c_contiguous
            r.counts = counts_out.ctypes.data
        r.records = records_out.ctypes.data if records_out.size else None
        r.record_capacity = records_out.size
        _check(self._L.rfb_scan_submit(self.ctx._h, self._h, C.byref(b), flags, C.byref(r)), self.ctx._h)
        self.ctx._inflight.append((r, data, records_out, counts_out))

    def wait(self):
        """Complete the oldest submitted batch: ScanResult over the arrays given to submit()."""
        done = C.POINTER(rfb_result)()
        _check(self._L.rfb_scan_wait(self.ctx._h, C.byref(done)), self.ctx._h)
        r, _data, records, counts = self.ctx._inflight.pop(0)
        return ScanResult(counts, records[: r.n_records], r, None)

    # -- device pointers in, device results out (for benchmarks: no PCIe in the timed region) --
    def scan_device(self, data_ptr, data_bytes, n_streams, n_steps, stride, counts_ptr=None, records_ptr=None,
                    record_capacity=0, flags=0, cuda_stream=None, offsets_ptr=None, steps_ptr=None,
                    stream_id_base=0, state_in_ptr=None, state_out_ptr=None, state_cap=0, pos_base=0):
        b = rfb_batch()
        b.data = data_ptr
        b.data_bytes = data_bytes
        b.n_streams = n_streams
        b.stride = stride
        b.n_steps = n_steps
        b.offsets = offsets_ptr
        b.steps = steps_ptr
        b.stream_id_base = stream_id_base
        b.pos_base = pos_base
        b.state_in = state_in_ptr
        b.state_out = state_out_ptr
        b.state_cap = state_cap
        r = rfb_result()
        r.counts = counts_ptr
        r.records = records_ptr
        r.record_capacity = record_capacity
        # cuda_stream: None -> the context's own stream; a cudaStream_t handle otherwise.  Handle 0 is the legacy
        # default stream (torch's default): pass cudaStreamLegacy (1) so that NULL keeps meaning "context stream".
        if cuda_stream is not None and int(cuda_stream) == 0:
            cuda_stream = 1
        _check(self._L.rfb_scan_device(self.ctx._h, self._h, C.byref(b), flags, cuda_stream, C.byref(r)), self.ctx._h)
        return r

    def calibrate(self, data, n_streams, n_steps, stride):
        """Order the start-DFA rows by the visit frequency measured on this HOST sample (rfb_nfa_calibrate).  The
        library does this by itself on the first large batch; calling it explicitly picks the sample and the moment."""
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        b = rfb_batch()
        b.data = data.ctypes.data if data.size else None
        b.data_bytes = data.size
        b.n_streams = n_streams
        b.stride = stride
        b.n_steps = n_steps
        _check(self._L.rfb_nfa_calibrate(self.ctx._h, self._h, C.byref(b)), self.ctx._h)

    def calibration(self):
        """(calibrated?, symbols of the sample, share of its start-DFA lookups served from shared memory)."""
        n, f = C.c_uint64(), C.c_double()
        rc = self._L.rfb_nfa_calibration(self._h, C.byref(n), C.byref(f))
        return bool(rc == 1), int(n.value), float(f.value)

    def fpga_cycles(self, lo, hi, trace_entries):
        """The testbench's "Total no. cycles" (testbench_BLK_Mem.sv:52,84) for an M-entry (lo, hi) trace pair."""
        lo = np.ascontiguousarray(lo, dtype=np.uint8)
        hi = np.ascontiguousarray(hi, dtype=np.uint8)
        assert lo.size >= trace_entries and hi.size >= trace_entries
        out = C.c_uint64()
        _check(self._L.rfb_fpga_cycles(self.ctx._h, self._h, lo.ctypes.data_as(C.POINTER(C.c_uint8)),
                                       hi.ctypes.data_as(C.POINTER(C.c_uint8)), trace_entries, C.byref(out)), self.ctx._h)
        return int(out.value)

    def collect(self, r):
        _check(self._L.rfb_scan_collect(self.ctx._h, C.byref(r)), self.ctx._h)
        return r


class Group:
    """N GPUs in one process (rfb_group_*): contiguous stream shards, one NCCL all-reduce of the per-state counts."""

    def __init__(self, device_ids):
        self._L = _lib.load()
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = C.c_void_p()
        _check(self._L.rfb_group_create(ids, len(device_ids), C.byref(h)))
        self._h = h
        self.size = int(self._L.rfb_group_size(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.rfb_group_destroy(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc:
            raise RfbError(rc, (self._L.rfb_group_last_error(self._h) or b"").decode())

    def nfa_from_entries(self, entries, n_states=-1):
        entries = np.ascontiguousarray(entries, dtype=np.uint32)
        h = C.c_void_p()
        self._check(self._L.rfb_group_nfa_from_entries(self._h, entries.ctypes.data_as(C.POINTER(C.c_uint32)), entries.size, n_states, C.byref(h)))
        return GroupNfa(self, h)


class GroupNfa:
    def __init__(self, group, handle):
        self.group, self._L, self._h = group, group._L, handle
        info = rfb_nfa_info()
        _check(self._L.rfb_nfa_get_info(self._L.rfb_group_nfa_member(self._h, 0), C.byref(info)))
        self.n_states = int(info.n_states)

    def close(self):
        if getattr(self, "_h", None):
            self._L.rfb_group_nfa_destroy(self._h)
            self._h = None

    def scan(self, data, n_streams, n_steps, stride, record_capacity=1 << 20, flags=SCAN_SORT_RECORDS, stream_id_base=0):
        data = np.ascontiguousarray(data, dtype=np.uint8).reshape(-1)
        b = rfb_batch()
        b.data = data.ctypes.data if data.size else None
        b.data_bytes = data.size
        b.n_streams, b.stride, b.n_steps, b.stream_id_base = n_streams, stride, n_steps, stream_id_base
        counts = np.zeros(self.n_states, dtype=np.uint64)
        records = np.empty(record_capacity, dtype=MATCH_DTYPE)
        r = rfb_result()
        r.counts = counts.ctypes.data
        r.records = records.ctypes.data if record_capacity else None
        r.record_capacity = record_capacity
        self.group._check(self._L.rfb_group_scan(self.group._h, self._h, C.byref(b), flags, C.byref(r)))
        return ScanResult(counts, records[: r.n_records], r, None)
