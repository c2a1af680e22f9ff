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


def make_batch_numpy(mix, lo, hi, n_streams, stream_len=1500, stride=1536, seed=0x5EED0001, first_stream=0):
    """Host batch as a (n_streams, stride) uint8 array."""
    out = np.zeros((n_streams, stride), dtype=np.uint8)
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
