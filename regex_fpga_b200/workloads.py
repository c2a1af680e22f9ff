"""Synthetic stream batches for BASELINE.json's configs.

A batch is n_streams x stream_len bytes laid out at `stride` bytes per stream (stride 1536 keeps every
stream 128-byte aligned; pad bytes are loaded but never scanned and are not counted in Gbit/s).

  W-mix (headline) : stream j is a stream_len-byte window of the shipped lo trace (j even) or hi trace
                     (j odd) starting at splitmix64(seed ^ j) mod (len(trace) - stream_len + 1)
  W-hi / W-lo      : all windows from one trace
  U                : i.i.d. uniform bytes (counter-based, splitmix64)
The same function bodies run on numpy (host, tests) and torch (device, bench) so both sides see the
same bytes.
"""
import numpy as np

MASK64 = (1 << 64) - 1


def splitmix64_np(x):
    """Vectorised splitmix64 finaliser over a uint64 array (wrapping arithmetic)."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = x + np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def window_offsets(seed, first_stream, n_streams, n_windows):
    j = np.arange(first_stream, first_stream + n_streams, dtype=np.uint64)
    return (splitmix64_np(np.uint64(seed) ^ j) % np.uint64(n_windows)).astype(np.int64)


def _source_select(mix, first_stream, n_streams):
    j = np.arange(first_stream, first_stream + n_streams)
    if mix == "wmix":
        return (j & 1).astype(bool)          # odd -> hi
    if mix == "whi":
        return np.ones(n_streams, dtype=bool)
    if mix == "wlo":
        return np.zeros(n_streams, dtype=bool)
    raise ValueError(mix)


SPLICE_SEGMENTS = 8


def splice_offsets(seed, first_stream, n_streams, n_windows):
    """wsplice: (n_streams, SPLICE_SEGMENTS) window offsets, one per segment of every stream."""
    j = np.arange(first_stream, first_stream + n_streams, dtype=np.uint64)[:, None]
    q = np.arange(SPLICE_SEGMENTS, dtype=np.uint64)[None, :]
    return (splitmix64_np(np.uint64(seed) ^ ((j << np.uint64(8)) | q) ^ np.uint64(0x5B11CE)) % np.uint64(n_windows)).astype(np.int64)


def make_batch_numpy(mix, lo, hi, n_streams, stream_len=1500, stride=1536, seed=0x5EED0001, first_stream=0):
    """Host batch as a (n_streams, stride) uint8 array.
    wsplice: every stream is SPLICE_SEGMENTS windows (of the lo trace for even streams, the hi trace for odd ones) from
    independent random offsets, back to back: the same byte statistics as W-mix but eight times the context switches, so
    the start DFA visits far more distinct rows (a batch of plain windows re-reads the same 2 x 200 KB of text)."""
    out = np.zeros((n_streams, stride), dtype=np.uint8)
    if mix == "wsplice":
        lo = np.asarray(lo, dtype=np.uint8)
        hi = np.asarray(hi, dtype=np.uint8)
        seg = -(-stream_len // SPLICE_SEGMENTS)
        n_windows = min(lo.size, hi.size) - seg + 1
        offs = splice_offsets(seed, first_stream, n_streams, n_windows)
        odd = (np.arange(first_stream, first_stream + n_streams) & 1).astype(bool)
        wl = np.lib.stride_tricks.sliding_window_view(lo, seg)
        wh = np.lib.stride_tricks.sliding_window_view(hi, seg)
        rows = np.where(odd[:, None, None], wh[offs], wl[offs]).reshape(n_streams, SPLICE_SEGMENTS * seg)
        out[:, :stream_len] = rows[:, :stream_len]
        return out
    if mix == "uniform":
        n_words = (stream_len + 7) // 8
        ctr = (np.arange(first_stream, first_stream + n_streams, dtype=np.uint64)[:, None] << np.uint64(32)) | \
            np.arange(n_words, dtype=np.uint64)[None, :]
        words = splitmix64_np(np.uint64(seed) ^ ctr)
        out[:, :stream_len] = words.view(np.uint8).reshape(n_streams, n_words * 8)[:, :stream_len]
        return out
    lo = np.asarray(lo, dtype=np.uint8)
    hi = np.asarray(hi, dtype=np.uint8)
    n_windows = min(lo.size, hi.size) - stream_len + 1
    offs = window_offsets(seed, first_stream, n_streams, n_windows)
    use_hi = _source_select(mix, first_stream, n_streams)
    wl = np.lib.stride_tricks.sliding_window_view(lo, stream_len)
    wh = np.lib.stride_tricks.sliding_window_view(hi, stream_len)
    idx_hi = np.nonzero(use_hi)[0]
    idx_lo = np.nonzero(~use_hi)[0]
    out[idx_hi, :stream_len] = wh[offs[idx_hi]]
    out[idx_lo, :stream_len] = wl[offs[idx_lo]]
    return out


def make_batch_torch(mix, lo, hi, n_streams, device, stream_len=1500, stride=1536, seed=0x5EED0001,
                     first_stream=0, chunk=1 << 16):
    """Device batch as a (n_streams, stride) uint8 torch tensor; byte-identical to make_batch_numpy."""
    import torch
    out = torch.zeros((n_streams, stride), dtype=torch.uint8, device=device)
    if mix == "wsplice":
        tl = torch.from_numpy(np.ascontiguousarray(lo, dtype=np.uint8)).to(device)
        th = torch.from_numpy(np.ascontiguousarray(hi, dtype=np.uint8)).to(device)
        seg = -(-stream_len // SPLICE_SEGMENTS)
        n_windows = min(tl.numel(), th.numel()) - seg + 1
        wl, wh = tl.unfold(0, seg, 1), th.unfold(0, seg, 1)
        for s0 in range(0, n_streams, chunk):
            n = min(chunk, n_streams - s0)
            offs = torch.from_numpy(splice_offsets(seed, first_stream + s0, n, n_windows)).to(device)
            odd = torch.from_numpy((np.arange(first_stream + s0, first_stream + s0 + n) & 1).astype(bool)).to(device)
            rows = torch.where(odd[:, None, None], wh[offs], wl[offs]).reshape(n, SPLICE_SEGMENTS * seg)
            out[s0:s0 + n, :stream_len] = rows[:, :stream_len]
        return out
    if mix == "uniform":
        for s0 in range(0, n_streams, chunk):
            n = min(chunk, n_streams - s0)
            part = make_batch_numpy(mix, None, None, n, stream_len, stride, seed, first_stream + s0)
            out[s0:s0 + n] = torch.from_numpy(part).to(device)
        return out
    tl = torch.from_numpy(np.ascontiguousarray(lo, dtype=np.uint8)).to(device)
    th = torch.from_numpy(np.ascontiguousarray(hi, dtype=np.uint8)).to(device)
    n_windows = min(tl.numel(), th.numel()) - stream_len + 1
    wl = tl.unfold(0, stream_len, 1)
    wh = th.unfold(0, stream_len, 1)
    for s0 in range(0, n_streams, chunk):
        n = min(chunk, n_streams - s0)
        offs = torch.from_numpy(window_offsets(seed, first_stream + s0, n, n_windows)).to(device)
        use_hi = torch.from_numpy(_source_select(mix, first_stream + s0, n)).to(device)
        rows = torch.where(use_hi[:, None], wh[offs], wl[offs])
        out[s0:s0 + n, :stream_len] = rows
    return out


# ---- BASELINE config 5: replicated large NFA and adversarial (high-activity) traces ---------------
def decode_rows(entries, n_states):
    """BRAM image -> (row_ptr, sym, tgt) numpy arrays (Design/FPGA.v:773,782,793,888-898)."""
    e = np.asarray(entries, dtype=np.uint32)
    rp = e[: n_states + 1].astype(np.int64)
    tr = e[n_states + 1: n_states + 1 + int(rp[n_states])]
    return rp, (tr >> np.uint32(24)).astype(np.int64), (tr & np.uint32(0xFFFFFF)).astype(np.int64)


def encode_rows(rp, sym, tgt):
    """(row_ptr, sym, tgt) -> BRAM image, zero-padded to whole 128-bit lines."""
    tr = ((np.asarray(sym, dtype=np.uint64) << np.uint64(24)) | np.asarray(tgt, dtype=np.uint64)).astype(np.uint32)
    e = np.concatenate([np.asarray(rp, dtype=np.uint32), tr])
    pad = (-e.size) % 4
    return np.concatenate([e, np.zeros(pad, np.uint32)])


def replicate_nfa(entries, n_states, copies):
    """`copies` disjoint replicas of an NFA behind ONE shared start state 0: replica r's state
    s >= 1 becomes 1 + r*(n-1) + (s-1); state 0's row is the concatenation of every replica's state-0 row.
    7 x snort_16 -> 66 592 states, 558 992 transitions (beyond the FPGA's own 16-bit rd_address)."""
    rp, sym, tgt = decode_rows(entries, n_states)
    n1 = n_states - 1
    assert not np.any(tgt == 0), "nothing may target state 0"
    rows_sym, rows_tgt, lens = [], [], []
    s0 = slice(int(rp[0]), int(rp[1]))
    rows_sym.append(np.concatenate([sym[s0]] * copies))
    rows_tgt.append(np.concatenate([tgt[s0] + r * n1 for r in range(copies)]))
    lens.append(rows_sym[0].size)
    body = slice(int(rp[1]), int(rp[n_states]))
    body_lens = np.diff(rp[1:])
    for r in range(copies):
        rows_sym.append(sym[body])
        rows_tgt.append(tgt[body] + r * n1)
        lens.extend(body_lens.tolist())
    new_rp = np.concatenate([[0], np.cumsum(np.asarray(lens, dtype=np.int64))])
    n_new = 1 + copies * n1
    assert new_rp.size == n_new + 1
    return encode_rows(new_rp, np.concatenate(rows_sym), np.concatenate(rows_tgt)), n_new


def adversarial_prefixes(entries, n_states, avoid=(0x0A, 0x0D), min_self=254, root=1):
    """Shortest symbol strings that drive the NFA from `root` (snort_16's global '.*' state) into each of its
    long-lived self-looping states without using the bytes that kill the [^\\n\\r]* ones.  Feeding several of
    them back to back accumulates persistent active states."""
    rp, sym, tgt = decode_rows(entries, n_states)
    avoid = set(avoid)
    self_cnt = np.zeros(n_states, np.int64)
    for s in range(n_states):
        self_cnt[s] = int(np.sum(tgt[rp[s]:rp[s + 1]] == s))
    goals = [s for s in range(n_states) if self_cnt[s] >= min_self and s != root]
    def bfs(banned):
        prev = {root: None}
        frontier = [root]
        while frontier:
            nxt = []
            for s in frontier:
                for j in range(int(rp[s]), int(rp[s + 1])):
                    t, c = int(tgt[j]), int(sym[j])
                    if t not in prev and c not in banned:
                        prev[t] = (s, c)
                        nxt.append(t)
            frontier = nxt
        return prev
    clean, anyb = bfs(avoid), bfs(set())   # prefer paths without the killer bytes, fall back to any path
    out = []
    for g in goals:
        prev = clean if g in clean else anyb
        if g not in prev:
            continue
        path, s = [], g
        while prev[s] is not None:
            s, c = prev[s]
            path.append(c)
        out.append(np.array(path[::-1], dtype=np.uint8))
    return out


def make_adversarial_numpy(entries, n_states, hi, n_streams, stream_len=1500, stride=1536, seed=0x5EED0005):
    """Streams that start with a random concatenation of adversarial prefixes (about half of the stream) and
    continue with a window of the hi trace."""
    pref = adversarial_prefixes(entries, n_states)
    rng = np.random.default_rng(seed)
    hi = np.asarray(hi, dtype=np.uint8)
    out = np.zeros((n_streams, stride), dtype=np.uint8)
    for s in range(n_streams):
        parts, total = [], 0
        while total < stream_len // 2 and pref:
            p = pref[int(rng.integers(len(pref)))]
            parts.append(p)
            total += p.size
        head = np.concatenate(parts)[: stream_len // 2] if parts else np.zeros(0, np.uint8)
        off = int(rng.integers(0, hi.size - stream_len))
        body = hi[off: off + stream_len - head.size]
        out[s, :stream_len] = np.concatenate([head, body])
    return out


def make_adversarial_torch(entries, n_states, hi, n_streams, device, stream_len=1500, stride=1536, seed=0x5EED0005,
                           n_base=4096, first_stream=0):
    """Device batch of adversarial streams: n_base distinct streams built on the host, then stream j is base
    stream splitmix64(seed ^ j) mod n_base (a stress workload, not a throughput headline)."""
    import torch
    base = torch.from_numpy(make_adversarial_numpy(entries, n_states, hi, n_base, stream_len, stride, seed)).to(device)
    pick = window_offsets(seed, first_stream, n_streams, n_base)
    return base[torch.from_numpy(pick).to(device)].contiguous()
