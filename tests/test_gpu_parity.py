"""GPU parity: the CUDA scan (through the C ABI) against the CPU oracle, bit-exact.

Everything here needs a B200 (`-m gpu`).  Records and counts must be identical to oracle B, which
tests/test_oracle.py pins against the cycle-level restatement of FPGA.v and SURVEY.md Appendix C.
"""
import numpy as np
import pytest

import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
from nfa_gen import build_entries, random_nfa, random_streams

pytestmark = pytest.mark.gpu

KERNELS = [("lane", R.SCAN_SORT_RECORDS), ("warp", R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP)]


def recs_tuple(r):
    return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))


def check_against_oracle(nfa, E, n_states, data2d, n_steps, flags, stride=None, cap=1 << 20):
    n_streams = data2d.shape[0]
    stride = data2d.shape[1] if stride is None else stride
    got = nfa.scan(data2d, n_streams, n_steps=n_steps, stride=stride, record_capacity=cap, flags=flags)
    want = O.b_scan_many(E, n_states, data2d, n_streams, stride, n_steps, cap=cap)
    assert got.n_matches == want["n_recs"]
    assert np.array_equal(got.counts, want["counts"])
    assert recs_tuple(got.records) == recs_tuple(want["recs"])
    assert got.n_symbols == n_streams * n_steps
    assert got.n_dropped == 0
    return got


@pytest.mark.parametrize("kernel,flags", KERNELS)
@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_testbench_run_matches_golden(gpu_ctx, snort, l7, expected, name, kernel, flags):
    """BASELINE configs 1-2: the committed testbench on the shipped trace pair (M = 200000)."""
    rs = snort if name == "snort_16" else l7
    nfa = gpu_ctx.nfa_from_entries(rs.entries)
    assert nfa.n_states == rs.n_states
    M = expected["tb_trace_entries"]
    data = np.stack([rs.lo[:M], rs.hi[:M]])          # stream 0 = lo = input_char, 1 = hi (TB:56-57)
    got = nfa.scan(data, 2, n_steps=R.tb_steps(M), stride=M, flags=flags)
    exp = expected["rulesets"][name]["tb"]
    for stream, key in ((0, "lo"), (1, "hi")):
        ev = got.records[got.records["stream"] == stream]
        assert [[int(p), int(s)] for p, s in zip(ev["pos"], ev["state"])] == exp[key]["events"]
    want_counts = np.zeros(rs.n_states, np.uint64)
    for key in ("lo", "hi"):
        for s, c in exp[key]["counts"].items():
            want_counts[int(s)] += c
    assert np.array_equal(got.counts, want_counts)
    assert got.n_matches == exp["lo"]["n_matches"] + exp["hi"]["n_matches"]
    if kernel == "lane":
        assert nfa.info["image_ok"] == 1 and got.n_rescanned == 0


def test_l7_full_files(gpu_ctx, l7, expected):
    """The l7 traces hold 262144 entries; the TB only consumes 200000."""
    nfa = gpu_ctx.nfa_from_entries(l7.entries)
    data = np.stack([l7.lo, l7.hi])
    got = nfa.scan(data, 2, n_steps=l7.lo.size - 1, stride=l7.lo.size)
    exp = expected["rulesets"]["l7_filter"]["full"]
    assert got.n_matches == exp["lo"]["n_matches"] + exp["hi"]["n_matches"] == 13
    ev = got.records[got.records["stream"] == 0]
    assert [[int(p), int(s)] for p, s in zip(ev["pos"], ev["state"])] == exp["lo"]["events"]


@pytest.mark.parametrize("kernel,flags", KERNELS)
@pytest.mark.parametrize("mix", ["wmix", "whi", "uniform"])
def test_packet_streams_match_oracle(gpu_ctx, snort, kernel, flags, mix):
    """BASELINE config 3 shape (1500-byte packet streams) on a subsample the oracle finishes quickly."""
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 3000 if kernel == "lane" else 600
    data = WL.make_batch_numpy(mix, snort.lo, snort.hi, n, 1500, 1536)
    got = check_against_oracle(nfa, snort.entries, snort.n_states, data, 1500, flags)
    if mix != "uniform":
        assert got.n_matches > 0


def test_l7_packet_streams(gpu_ctx, l7):
    nfa = gpu_ctx.nfa_from_entries(l7.entries)
    data = WL.make_batch_numpy("wmix", l7.lo, l7.hi, 2000, 1500, 1536)
    check_against_oracle(nfa, l7.entries, l7.n_states, data, 1500, R.SCAN_SORT_RECORDS)


@pytest.mark.parametrize("seed", range(25))
def test_random_nfas_both_kernels(gpu_ctx, seed):
    rng = np.random.default_rng(7000 + seed)
    (E, n), syms = random_nfa(rng, n_states=int(rng.integers(2, 400)), alphabet=int(rng.integers(2, 30)),
                              p_sticky=float(rng.choice([0.0, 0.1, 0.4])), max_fanout=int(rng.integers(1, 4)))
    nfa = gpu_ctx.nfa_from_entries(E, n)
    length = int(rng.integers(1, 300))
    data = random_streams(rng, syms, int(rng.integers(1, 200)), length)
    for _, flags in KERNELS:
        check_against_oracle(nfa, E, n, data, length, flags)


@pytest.mark.parametrize("budget", ["0", "24", "200", "16384"])
@pytest.mark.parametrize("seed", range(6))
def test_start_dfa_unanchored_random_nfas(gpu_ctx, seed, budget, monkeypatch):
    """Unanchored NFAs (state 0 -> ".*" state on every symbol): the lane kernel follows the ordinary successors of
    the always-active state with a start DFA.  RFB_DFA_STATES = 0 disables it, small budgets force the failure-link
    fallback (transitions beyond the budget continue from a shorter history and insert the rest explicitly).
    Every variant must equal oracle B, one-shot and split into two resumed calls."""
    monkeypatch.setenv("RFB_DFA_STATES", budget)
    rng = np.random.default_rng(9100 + seed)
    (E, n), syms = random_nfa(rng, n_states=int(rng.integers(12, 300)), alphabet=int(rng.integers(3, 12)),
                              p_sticky=float(rng.choice([0.0, 0.05, 0.2])), max_fanout=int(rng.integers(1, 4)),
                              unanchored=True)
    nfa = gpu_ctx.nfa_from_entries(E, n)
    assert nfa.info["image_ok"] == 1
    length = int(rng.integers(40, 260))
    data = random_streams(rng, syms, int(rng.integers(8, 120)), length)
    whole = check_against_oracle(nfa, E, n, data, length, R.SCAN_SORT_RECORDS)
    cut = length // 2
    a = nfa.scan(np.ascontiguousarray(data[:, :cut]), data.shape[0], n_steps=cut, stride=cut, want_state=True, state_cap=255)
    b = nfa.scan(np.ascontiguousarray(data[:, cut:]), data.shape[0], n_steps=length - cut, stride=length - cut,
                 state_in=a.state, pos_base=cut)
    if not np.any(a.state[:, 0] == R.STATE_OVERFLOW):
        assert sorted(recs_tuple(a.records) + recs_tuple(b.records)) == recs_tuple(whole.records)
        for s in range(min(8, data.shape[0])):     # exported sets hold every member exactly once
            ids = a.state[s, 1:1 + a.state[s, 0]]
            assert len(set(ids.tolist())) == len(ids)


@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_start_dfa_budget_on_shipped_rulesets(gpu_ctx, snort, l7, name, monkeypatch):
    """The shipped images with a start DFA cut at 300 states (most rows resolve through failure links)."""
    rs = snort if name == "snort_16" else l7
    monkeypatch.setenv("RFB_DFA_STATES", "300")
    nfa = gpu_ctx.nfa_from_entries(rs.entries)
    data = WL.make_batch_numpy("wmix", rs.lo, rs.hi, 64, 1500, 1536, seed=0x5EED0044)
    check_against_oracle(nfa, rs.entries, rs.n_states, data, 1500, R.SCAN_SORT_RECORDS, stride=1536)


def test_high_activity_overflows_to_warp_kernel(gpu_ctx):
    """More simultaneously active transient states than the lane kernel's ring holds: those streams
    must be re-run by the general kernel and still be bit-exact, with no double reports."""
    n = 64
    rows = [[(1, s) for s in range(1, n - 1)]]                       # 0 fans out to everything
    for s in range(1, n - 1):
        rows.append([(1, s + 1 if s + 1 < n - 1 else 1), (2, n - 1), (1, n - 1)])
    rows.append([])                                                  # accept
    E, ns = build_entries(rows)
    nfa = gpu_ctx.nfa_from_entries(E, ns)
    rng = np.random.default_rng(5)
    data = rng.choice(np.array([1, 1, 1, 2, 3], np.uint8), size=(300, 200))
    data[::3] = 3                                                    # a third of the streams stay quiet
    got = check_against_oracle(nfa, E, ns, data, 200, R.SCAN_SORT_RECORDS)
    assert got.n_rescanned > 0 and got.n_rescanned <= 200
    check_against_oracle(nfa, E, ns, data, 200, R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP)


def test_ragged_offsets_and_empty_streams(gpu_ctx, snort):
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    rng = np.random.default_rng(11)
    n = 500
    steps = rng.integers(0, 700, size=n).astype(np.uint32)
    steps[::17] = 0                                                  # empty streams
    steps[5] = 1
    offsets = np.zeros(n, np.uint64)
    pos = 3                                                          # deliberately unaligned starts
    for s in range(n):
        offsets[s] = pos
        pos += int(steps[s]) + int(rng.integers(0, 5))
    buf = np.zeros(pos + 16, np.uint8)
    src = np.concatenate([snort.hi, snort.lo])
    for s in range(n):
        o = int(rng.integers(0, src.size - 700))
        buf[int(offsets[s]): int(offsets[s]) + int(steps[s])] = src[o: o + int(steps[s])]
    for _, flags in KERNELS:
        got = nfa.scan(buf, n, offsets=offsets, steps=steps, flags=flags, stream_id_base=1000)
        want_counts = np.zeros(snort.n_states, np.uint64)
        want = []
        for s in range(n):
            b = O.b_scan(snort.entries, snort.n_states, buf[int(offsets[s]):], int(steps[s]), stream_id=1000 + s)
            want_counts += b["counts"]
            want += recs_tuple(b["recs"])
        assert np.array_equal(got.counts, want_counts)
        assert recs_tuple(got.records) == want
        assert got.n_symbols == int(steps.sum())
    # zero streams / zero steps are legal and report nothing
    got = nfa.scan(np.zeros(16, np.uint8), 0, n_steps=10, stride=10)
    assert got.n_matches == 0 and got.n_symbols == 0
    got = nfa.scan(np.zeros(64, np.uint8), 4, n_steps=0, stride=16)
    assert got.n_matches == 0 and got.n_symbols == 0


def test_record_buffer_overflow_is_counted(gpu_ctx, snort):
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    data = WL.make_batch_numpy("whi", snort.lo, snort.hi, 512, 1500, 1536)
    full = nfa.scan(data, 512, n_steps=1500, stride=1536)
    assert full.n_matches > 200
    small = nfa.scan(data, 512, n_steps=1500, stride=1536, record_capacity=100)
    assert small.n_matches == full.n_matches and small.n_records == 100
    assert small.n_dropped == full.n_matches - 100
    assert np.array_equal(small.counts, full.counts)                 # counts never depend on capacity
    allrec = set(recs_tuple(full.records))
    assert all(r in allrec for r in recs_tuple(small.records))
    none = nfa.scan(data, 512, n_steps=1500, stride=1536, record_capacity=0)
    assert none.n_records == 0 and none.n_dropped == full.n_matches
    assert np.array_equal(none.counts, full.counts)


def test_accept_start_state_and_tiny_nfas(gpu_ctx):
    E, n = build_entries([[]])                                       # size 1: state 0 accepts at step 0 only
    nfa = gpu_ctx.nfa_from_entries(E, n)
    data = np.zeros((3, 8), np.uint8)
    for _, flags in KERNELS:
        got = nfa.scan(data, 3, n_steps=8, stride=8, flags=flags)
        assert recs_tuple(got.records) == [(0, 0, 0), (1, 0, 0), (2, 0, 0)]
    E, n = build_entries([[(7, 0), (7, 1)], []])                     # 0 loops on 7 and feeds accept 1
    nfa = gpu_ctx.nfa_from_entries(E, n)
    d = np.full((2, 2000), 7, np.uint8)
    d[1, 1000] = 8
    for _, flags in KERNELS:
        got = check_against_oracle(nfa, E, n, d, 2000, flags)
        assert got.counts[1] == 1999 + 1000


def test_device_resident_scan_matches_host_scan(gpu_ctx, snort):
    """rfb_scan_device (the entry point bench.py times) against rfb_scan on the same batch."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 4096
    host = WL.make_batch_numpy("wmix", snort.lo, snort.hi, n, 1500, 1536)
    dev = WL.make_batch_torch("wmix", snort.lo, snort.hi, n, "cuda:0", 1500, 1536)
    assert np.array_equal(dev.cpu().numpy(), host)                   # generators agree byte for byte
    counts = torch.zeros(snort.n_states, dtype=torch.int64, device="cuda:0")
    cap = 1 << 16
    recs = torch.zeros(cap * 3, dtype=torch.int32, device="cuda:0")
    r = nfa.scan_device(dev.data_ptr(), dev.numel(), n, 1500, 1536, counts.data_ptr(), recs.data_ptr(), cap,
                        cuda_stream=torch.cuda.current_stream().cuda_stream)
    ref = nfa.scan(host, n, n_steps=1500, stride=1536)
    assert r.n_matches == ref.n_matches and r.n_symbols == n * 1500
    assert np.array_equal(counts.cpu().numpy().astype(np.uint64), ref.counts)
    got = recs.cpu().numpy().view(np.uint32).reshape(-1, 3)[: r.n_records]
    got = sorted(map(tuple, got.tolist()))
    assert got == recs_tuple(ref.records)


@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_fpga_cycles_and_tb_report(gpu_ctx, snort, l7, expected, name, tmp_path):
    """the testbench's printout from GPU results -- per-state counters of both streams and
    `Total no. cycles` (closed-form cycle model of FPGA.v evaluated on the GPU) -- against the golden values the
    cycle-level oracle produced (2 188 184 738 / 617 518 104 cycles)."""
    from regex_fpga_b200 import tbreport
    rs = snort if name == "snort_16" else l7
    nfa = gpu_ctx.nfa_from_entries(rs.entries)
    exp = expected["rulesets"][name]["tb"]
    assert nfa.fpga_cycles(rs.lo, rs.hi, 2000) == exp["cycles_first_2000_entries"]
    assert nfa.fpga_cycles(rs.lo, rs.hi, 1) == 1                      # reset edge only
    lines, counters = tbreport.tb_report(nfa, rs.lo, rs.hi, expected["tb_trace_entries"])
    assert lines[-1] == f"Total no. cycles: {exp['total_cycles']}"
    for key, mc, label in (("lo", counters[0], "match_count"), ("hi", counters[1], "match_count_2")):
        want = {int(s): c & 0x3FF for s, c in exp[key]["counts"].items()}
        assert {int(p): int(mc[p]) for p in np.nonzero(mc)[0]} == want
        mine = [ln for ln in lines if ln.startswith(label + "[")]
        assert mine == [f"{label}[{p}] = {want[p]}" for p in sorted(want, reverse=True)]
    # the CLI end to end, through the reference's own file formats
    coe, lo, hi = tmp_path / "n.coe", tmp_path / "lo.mem", tmp_path / "hi.mem"
    R.coe_write(coe, rs.entries, 1)
    R.trace_write_mem(lo, rs.lo[:3000])
    R.trace_write_mem(hi, rs.hi[:3000])
    import io, contextlib
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        tbreport.main(["--coe", str(coe), "--lo", str(lo), "--hi", str(hi), "--entries", "3000"])
    out = buf.getvalue().strip().split("\n")
    assert out[-1] == f"Total no. cycles: {O.cycle_model(rs.entries, rs.n_states, rs.lo, rs.hi, 3000)}"


@pytest.mark.parametrize("kernel,flags", KERNELS)
def test_resumable_scan_equals_one_shot(gpu_ctx, snort, l7, kernel, flags):
    """the design is a streaming device (input_char_flag handshake, FPGA.v:125,740,763) whose only
    carry-over state is the active set.  Cutting streams at arbitrary points and resuming from the exported sets
    must reproduce the one-shot scan exactly: same records (pos_base added), same counts."""
    rng = np.random.default_rng(3)
    for rs in (snort, l7):
        nfa = gpu_ctx.nfa_from_entries(rs.entries)
        n, L = 64, 4000
        data = np.stack([(rs.hi if s & 1 else rs.lo)[s * 700: s * 700 + L] for s in range(n)])
        whole = nfa.scan(data, n, n_steps=L, stride=L, flags=flags)
        cuts = [0] + sorted(rng.choice(np.arange(1, L), size=5, replace=False).tolist()) + [L]
        cuts.insert(3, cuts[3])                                     # an empty chunk in the middle
        state, counts, recs = None, np.zeros(rs.n_states, np.uint64), []
        for a, b in zip(cuts[:-1], cuts[1:]):
            part = nfa.scan(np.ascontiguousarray(data[:, a:b]) if b > a else np.zeros((n, 1), np.uint8), n,
                            n_steps=b - a, stride=max(b - a, 1), flags=flags, state_in=state, want_state=True,
                            pos_base=a, state_cap=127)
            assert not np.any(part.state[:, 0] == R.STATE_OVERFLOW)
            state = part.state
            counts += part.counts
            recs += recs_tuple(part.records)
        assert np.array_equal(counts, whole.counts)
        assert sorted(recs) == recs_tuple(whole.records)
        # the exported set is S_{n_steps} in the reference's state numbering: compare with the oracle's next sets
        for s in (0, 1, 33):
            want = set()
            cur = {0}
            rp = rs.entries[: rs.n_states + 1]
            tr = rs.entries[rs.n_states + 1:]
            for k in range(cuts[1]):
                c = data[s, k]
                nxt = set()
                for st_ in cur:
                    row = tr[rp[st_]: rp[st_ + 1]]
                    nxt.update((row[(row >> 24) == c] & 0xFFFFFF).tolist())
                cur = nxt
            want = cur
            first = nfa.scan(np.ascontiguousarray(data[:, : cuts[1]]), n, n_steps=cuts[1], stride=cuts[1], flags=flags,
                             want_state=True, state_cap=127).state
            assert set(first[s, 1: 1 + first[s, 0]].tolist()) == want


def test_resume_with_more_states_than_the_ring_holds(gpu_ctx):
    """A resumed set larger than the lane kernel's ring goes to the general kernel from step 0; an exported set
    larger than state_cap is flagged, not truncated."""
    n = 80
    rows = [[(1, s) for s in range(1, n - 1)]]
    for s in range(1, n - 1):
        rows.append([(1, s + 1 if s + 1 < n - 1 else 1), (2, n - 1)])
    rows.append([])
    E, ns = build_entries(rows)
    nfa = gpu_ctx.nfa_from_entries(E, ns)
    d = np.ones((8, 40), np.uint8)
    d[:, 30] = 2
    whole = nfa.scan(d, 8, n_steps=40, stride=40)
    a = nfa.scan(np.ascontiguousarray(d[:, :10]), 8, n_steps=10, stride=10, want_state=True, state_cap=100)
    assert np.all(a.state[:, 0] == n - 2)                           # 78 active states: more than the ring's 31..63
    b = nfa.scan(np.ascontiguousarray(d[:, 10:]), 8, n_steps=30, stride=30, state_in=a.state, pos_base=10)
    assert sorted(recs_tuple(a.records) + recs_tuple(b.records)) == recs_tuple(whole.records)
    assert b.n_rescanned == 8
    small = nfa.scan(np.ascontiguousarray(d[:, :10]), 8, n_steps=10, stride=10, want_state=True, state_cap=16)
    assert np.all(small.state[:, 0] == R.STATE_OVERFLOW)


@pytest.mark.parametrize("seed", [500337, 500008, 500014, 500045])
def test_resume_when_the_ring_overflows_in_the_last_step(gpu_ctx, seed):
    """Regression (found by tools/dev/stress_parity.py): a stream whose ring filled up while S_{n_steps} was being
    built had nothing left to report, so the lane kernel did not hand it over -- and nobody wrote its state_out row.
    Now the general kernel re-runs it for the state, and rows start out as the overflow mark."""
    rng = np.random.default_rng(seed)
    rng.choice([0, 12, 40, 300, 16384])                              # same draw order as the soak tool
    (E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 500)), alphabet=int(rng.integers(2, 20)),
                              p_sticky=float(rng.choice([0.0, 0.05, 0.2, 0.4])), p_accept=float(rng.choice([0.05, 0.15, 0.3])),
                              max_fanout=int(rng.integers(1, 5)), unanchored=bool(rng.integers(0, 4)))
    L = int(rng.integers(2, 400)); ns = int(rng.integers(1, 150))
    data = random_streams(rng, syms, ns, L, p_alpha=float(rng.choice([0.6, 0.85, 0.97])))
    nfa = gpu_ctx.nfa_from_entries(E, n)
    want = O.b_scan_many(E, n, data, ns, L, L, cap=1 << 22)
    cut = L // 2
    a = nfa.scan(np.ascontiguousarray(data[:, :cut]), ns, n_steps=cut, stride=cut, want_state=True, state_cap=255, record_capacity=1 << 22)
    assert a.n_rescanned > 0 and not np.any(a.state[:, 0] == R.STATE_OVERFLOW)
    b = nfa.scan(np.ascontiguousarray(data[:, cut:]), ns, n_steps=L - cut, stride=L - cut, state_in=a.state, pos_base=cut,
                 record_capacity=1 << 22)
    assert sorted(recs_tuple(a.records) + recs_tuple(b.records)) == recs_tuple(want["recs"])


def test_state_rows_that_cannot_be_resumed(gpu_ctx, snort):
    """An overflow mark or a foreign id in state_in: rfb_scan refuses host rows; the device-pointer variant ignores
    such entries instead of indexing out of bounds.  A cut NFA whose set outgrows state_cap says so in every part."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    data = WL.make_batch_numpy("whi", snort.lo, snort.hi, 4, 200, 256, seed=3)
    st = np.zeros((4, 8), np.uint32)
    st[:, 0] = 1
    st[2, 0] = R.STATE_OVERFLOW
    with pytest.raises(R.RfbError, match="overflow mark"):
        nfa.scan(data, 4, n_steps=200, stride=256, state_in=st)
    st[2, 0] = 1
    st[1, 1] = snort.n_states + 5
    with pytest.raises(R.RfbError, match="not a state"):
        nfa.scan(data, 4, n_steps=200, stride=256, state_in=st)
    # device rows are not readable by the host: bad entries are skipped (stream 1 resumes from the empty set)
    st[1, 1] = 0xFFFFFF00
    st[3, 0] = R.STATE_OVERFLOW
    d_st = torch.from_numpy(st.view(np.int32)).to("cuda:0")
    d_data = torch.from_numpy(data).to("cuda:0")
    counts = torch.zeros(snort.n_states, dtype=torch.int64, device="cuda:0")
    for flags in (0, R.SCAN_FORCE_WARP):
        r = nfa.scan_device(d_data.data_ptr(), d_data.numel(), 4, 200, 256, counts.data_ptr(), None, 0, flags=flags,
                            cuda_stream=torch.cuda.current_stream().cuda_stream, state_in_ptr=d_st.data_ptr(), state_cap=7)
        assert r.n_symbols == 800
    # 3 replicas behind one start state: every replica's ".*" state is active after one symbol, 3+ ids > state_cap 2
    E, n = WL.replicate_nfa(snort.entries, snort.n_states, 3)
    big = gpu_ctx.nfa_from_entries(E, n)
    assert big.info["n_parts"] >= 2
    a = big.scan(data, 4, n_steps=50, stride=256, want_state=True, state_cap=2)
    assert np.all(a.state[:, 0] == R.STATE_OVERFLOW)


def test_config5_replicated_large_nfa_adversarial(gpu_ctx, snort):
    """BASELINE config 5: 7 x snort_16 behind one start state (66 592 states, beyond the FPGA's own 16-bit
    rd_address) with adversarial high-activity streams and hi-trace windows.  The tables of the whole NFA do not fit
    one SM, so the library cuts it along its connected components into 7 parts and scans the batch once per part
    (lane kernel); FORCE_WARP runs the general kernel on the same parts.  Bit-exact vs oracle B on the FULL NFA."""
    E7, n7 = WL.replicate_nfa(snort.entries, snort.n_states, 7)
    assert n7 == 66592 and int(E7[n7]) == 558992
    nfa = gpu_ctx.nfa_from_entries(E7, n7)
    assert nfa.info["n_parts"] == 7 and nfa.info["image_ok"] == 1
    adv = WL.make_adversarial_numpy(snort.entries, snort.n_states, snort.hi, 96)
    whi = WL.make_batch_numpy("whi", snort.lo, snort.hi, 96, 1500, 1536, seed=0x5EED0005)
    data = np.concatenate([adv, whi])
    for _, flags in KERNELS:
        got = check_against_oracle(nfa, E7, n7, data, 1500, flags)
        assert got.n_matches > 100
    # every replica sees the same bytes, so a match on state s of replica 0 appears on all 7 replicas
    one = gpu_ctx.nfa_from_entries(snort.entries).scan(data, data.shape[0], n_steps=1500, stride=1536)
    assert got.n_matches == 7 * one.n_matches
    per = got.counts[1:].reshape(7, snort.n_states - 1)
    assert all(np.array_equal(per[0], per[r]) for r in range(1, 7))
    assert np.array_equal(per[0], one.counts[1:])
    # ragged lengths: symbols are counted once, not once per part
    steps = np.full(data.shape[0], 1500, np.uint32)
    steps[::3] = 100
    rag = nfa.scan(data, data.shape[0], stride=1536, steps=steps)
    assert rag.n_symbols == int(steps.sum())
    # resumable scans across the parts: every part imports its own members and appends them on export
    sub = np.ascontiguousarray(data[:24])
    whole = nfa.scan(sub, 24, n_steps=1500, stride=1536)
    a = nfa.scan(sub, 24, n_steps=700, stride=1536, want_state=True, state_cap=255)
    assert not np.any(a.state[:, 0] == R.STATE_OVERFLOW) and a.state[:, 0].max() > 7
    b = nfa.scan(np.ascontiguousarray(sub[:, 700:]), 24, n_steps=800, stride=1536 - 700, state_in=a.state, pos_base=700)
    assert sorted(recs_tuple(a.records) + recs_tuple(b.records)) == recs_tuple(whole.records)


def test_unsplittable_large_nfa_uses_general_kernel(gpu_ctx):
    """One connected component too large for the lane tables: a single part on the general kernel."""
    n = 40000
    rows = [[(1, 1)]] + [[(1, (s % (n - 1)) + 1), (2, ((s * 7) % (n - 1)) + 1)] for s in range(1, n - 1)] + [[]]
    rows[5] = [(1, 6), (3, n - 1)]
    E, ns = build_entries(rows)
    nfa = gpu_ctx.nfa_from_entries(E, ns)
    assert nfa.info["n_parts"] == 1 and nfa.info["image_ok"] == 0
    rng = np.random.default_rng(9)
    data = rng.choice(np.array([1, 1, 2, 3], np.uint8), size=(64, 120))
    check_against_oracle(nfa, E, ns, data, 120, R.SCAN_SORT_RECORDS)


def test_host_path_chunked_overlap_matches_device_path(gpu_ctx, snort):
    """rfb_scan copies batches >= 128 MB in chunks on a second stream while ONE kernel launch consumes them
    (warps wait on a per-chunk arrival counter).  Counts and records must equal the device-resident scan,
    also with per-stream lengths and a non-default stream_id_base."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 110000                                                        # 169 MB at stride 1536 -> 2 chunks; not a multiple of 32
    dev = WL.make_batch_torch("wmix", snort.lo, snort.hi, n, "cuda:0", 1500, 1536)
    host = dev.cpu().numpy()
    counts = torch.zeros(snort.n_states, dtype=torch.int64, device="cuda:0")
    cap = 1 << 19
    recs = torch.zeros(cap * 3, dtype=torch.int32, device="cuda:0")
    r = nfa.scan_device(dev.data_ptr(), dev.numel(), n, 1500, 1536, counts.data_ptr(), recs.data_ptr(), cap,
                        cuda_stream=torch.cuda.current_stream().cuda_stream, stream_id_base=7)
    want = sorted(map(tuple, recs.cpu().numpy().view(np.uint32).reshape(-1, 3)[: r.n_records].tolist()))
    got = nfa.scan(host, n, n_steps=1500, stride=1536, record_capacity=cap, stream_id_base=7)
    assert got.n_matches == r.n_matches and got.n_symbols == n * 1500
    assert np.array_equal(got.counts, counts.cpu().numpy().astype(np.uint64))
    assert recs_tuple(got.records) == want
    # caller-owned (pinned) result arrays, as bench.py's e2e leg passes them
    prec = torch.empty(cap * 12, dtype=torch.uint8, pin_memory=True).numpy().view(R.engine.MATCH_DTYPE)
    pcnt = torch.empty(snort.n_states, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    own = nfa.scan(host, n, n_steps=1500, stride=1536, records_out=prec, counts_out=pcnt, stream_id_base=7)
    assert own.records.base is not None and recs_tuple(own.records) == want and np.array_equal(pcnt, got.counts)
    steps = np.full(n, 1500, np.uint32)
    steps[::5] = 700
    steps[3::7] = 0
    ragged = nfa.scan(host, n, stride=1536, steps=steps, record_capacity=cap)
    sub = np.arange(0, n, 997)
    for s in sub[:60]:
        b = O.b_scan(snort.entries, snort.n_states, host[s], int(steps[s]), stream_id=int(s))
        mine = ragged.records[ragged.records["stream"] == s]
        assert recs_tuple(mine) == recs_tuple(b["recs"])
    assert ragged.n_symbols == int(steps.sum())


def test_pipelined_host_path_matches_rfb_scan(gpu_ctx, snort):
    """rfb_scan_submit / rfb_scan_wait: two host batches in flight (the second batch's H2D copy overlaps the first
    one's tail, sort and D2H).  Results must equal rfb_scan's, batch by batch, in submission order; a third submit
    without a wait, a wait with nothing in flight and ragged batches are refused."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 100000                                                        # 154 MB: chunked copy + gated kernel
    batches = [WL.make_batch_torch("wmix", snort.lo, snort.hi, n, "cuda:0", 1500, 1536, seed=0x5EED0100 + i).cpu().pin_memory()
               for i in range(3)]
    cap = 1 << 19
    want = [nfa.scan(b.numpy(), n, n_steps=1500, stride=1536, record_capacity=cap, stream_id_base=10 * i) for i, b in enumerate(batches)]
    recs = [torch.empty(cap * 12, dtype=torch.uint8, pin_memory=True).numpy().view(R.engine.MATCH_DTYPE) for _ in range(2)]
    cnts = [torch.empty(snort.n_states, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64) for _ in range(2)]
    got = []
    for i, b in enumerate(batches):
        nfa.submit(b.numpy(), n, 1500, 1536, recs[i & 1], cnts[i & 1], stream_id_base=10 * i)
        if i >= 1:
            r = nfa.wait()
            got.append((recs_tuple(r.records), r.counts.copy(), r.n_matches, r.n_symbols))
    r = nfa.wait()
    got.append((recs_tuple(r.records), r.counts.copy(), r.n_matches, r.n_symbols))
    for w, (g_recs, g_counts, g_m, g_sym) in zip(want, got):
        assert g_m == w.n_matches and g_sym == n * 1500 and np.array_equal(g_counts, w.counts) and g_recs == recs_tuple(w.records)
    with pytest.raises(R.RfbError, match="no batch"):
        nfa.wait()
    small = batches[0].numpy()[:64]
    nfa.submit(small, 64, 1500, 1536, recs[0], cnts[0])
    nfa.submit(small, 64, 1500, 1536, recs[1], cnts[1])
    with pytest.raises(R.RfbError, match="two batches"):
        nfa.submit(small, 64, 1500, 1536, recs[0], cnts[0])
    a, b2 = nfa.wait(), nfa.wait()
    assert a.n_matches == b2.n_matches and recs_tuple(a.records) == recs_tuple(b2.records)


def test_device_side_record_sort(gpu_ctx, snort):
    """RFB_SCAN_SORT_RECORDS on the device path: a stable LSD radix sort over the 12-byte records must give exactly
    the canonical (stream, pos, state) order, also with a large stream_id_base / pos_base (all key bytes in use)."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 60000
    dev = WL.make_batch_torch("whi", snort.lo, snort.hi, n, "cuda:0", 1500, 1536)
    cap = 1 << 19
    for sbase, pbase in ((0, 0), (0xF0000000, 0x7F000000)):
        recs = torch.zeros(cap * 3, dtype=torch.int32, device="cuda:0")
        r = nfa.scan_device(dev.data_ptr(), dev.numel(), n, 1500, 1536, None, recs.data_ptr(), cap, flags=R.SCAN_SORT_RECORDS,
                            cuda_stream=torch.cuda.current_stream().cuda_stream, stream_id_base=sbase, pos_base=pbase)
        assert r.n_records == r.n_matches > 100000
        got = recs.cpu().numpy().view(np.uint32).reshape(-1, 3)[: r.n_records]
        order = np.lexsort((got[:, 2], got[:, 1], got[:, 0]))
        assert np.array_equal(order, np.arange(r.n_records))           # already sorted, ties impossible (records are unique)
        assert int(got[:, 0].min()) >= sbase and int(got[:, 1].min()) >= pbase
    unsorted = torch.zeros(cap * 3, dtype=torch.int32, device="cuda:0")
    r2 = nfa.scan_device(dev.data_ptr(), dev.numel(), n, 1500, 1536, None, unsorted.data_ptr(), cap,
                         cuda_stream=torch.cuda.current_stream().cuda_stream, stream_id_base=0xF0000000, pos_base=0x7F000000)
    u = unsorted.cpu().numpy().view(np.uint32).reshape(-1, 3)[: r2.n_records]
    assert sorted(map(tuple, u.tolist())) == list(map(tuple, got.tolist()))


def test_large_batch_properties(gpu_ctx, snort):
    """Full-size-shaped batch (256K streams here; bench runs 1M): size-independent properties --
    counts are additive over any partition of the streams, independent of stream order, and the lane
    and warp kernels agree."""
    import torch
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 1 << 18
    dev = WL.make_batch_torch("wmix", snort.lo, snort.hi, n, "cuda:0", 1500, 1536)

    def run(t, flags=0):
        c = torch.zeros(snort.n_states, dtype=torch.int64, device="cuda:0")
        r = nfa.scan_device(t.data_ptr(), t.numel(), t.shape[0], 1500, 1536, c.data_ptr(), None, 0, flags=flags,
                            cuda_stream=torch.cuda.current_stream().cuda_stream)   # order after torch's own work
        return c.cpu().numpy(), r.n_matches

    c_all, m_all = run(dev)
    c_a, m_a = run(dev[: n // 3].contiguous())
    c_b, m_b = run(dev[n // 3:].contiguous())
    assert m_all == m_a + m_b and np.array_equal(c_all, c_a + c_b) and m_all == int(c_all.sum())
    perm = torch.randperm(n, device="cuda:0", generator=torch.Generator("cuda:0").manual_seed(1))
    c_p, m_p = run(dev[perm].contiguous())
    assert m_p == m_all and np.array_equal(c_p, c_all)
    sub = dev[: 1 << 14].contiguous()
    c_l, _ = run(sub)
    c_w, _ = run(sub, R.SCAN_FORCE_WARP)
    assert np.array_equal(c_l, c_w)
    # and a strided subsample against the oracle
    idx = np.arange(0, n, n // 512)[:512]
    host = dev[torch.from_numpy(idx).to("cuda:0")].cpu().numpy()
    want = O.b_scan_many(snort.entries, snort.n_states, host, 512, 1536, 1500, want_recs=False)
    c_s, _ = run(torch.from_numpy(host).to("cuda:0"))
    assert np.array_equal(c_s.astype(np.uint64), want["counts"])
    # the full per-state count vector of a 32 768-stream block against the multi-threaded oracle, and
    # the records of the first 4 096 + 4 096 strided streams
    blk = dev[: 1 << 15].contiguous()
    c_blk, m_blk = run(blk)
    want = O.b_scan_many(snort.entries, snort.n_states, blk.cpu().numpy(), 1 << 15, 1536, 1500, want_recs=False)
    assert np.array_equal(c_blk.astype(np.uint64), want["counts"]) and m_blk == int(want["counts"].sum())
    idx = np.concatenate([np.arange(4096), np.arange(4096, n, (n - 4096) // 4096)[:4096]])
    host = dev[torch.from_numpy(idx).to("cuda:0")].cpu().numpy()
    got = nfa.scan(host, host.shape[0], n_steps=1500, stride=1536, record_capacity=1 << 20)
    want = O.b_scan_many(snort.entries, snort.n_states, host, host.shape[0], 1536, 1500, cap=1 << 20)
    assert got.n_dropped == 0 and recs_tuple(got.records) == recs_tuple(want["recs"])


def test_describe_reports_the_start_dfa(gpu_ctx, snort, l7, monkeypatch):
    text = gpu_ctx.nfa_from_entries(snort.entries).describe()       # sticky states moved into a still complete DFA
    assert "kernel lane" in text and "dfa_beyond_budget 0 " in text and "dfa_absorbed_sticky 0" not in text
    monkeypatch.setenv("RFB_DFA_ABSORB", "0")
    for rs, states in ((snort, 8495), (l7, 1763)):
        text = gpu_ctx.nfa_from_entries(rs.entries).describe()
        assert "kernel lane" in text and f"dfa_states {states} " in text and "dfa_beyond_budget 0 " in text   # complete DFAs
    monkeypatch.setenv("RFB_DFA_STATES", "300")
    text = gpu_ctx.nfa_from_entries(snort.entries).describe()
    assert "dfa_states 300 " in text and "dfa_beyond_budget 0 " not in text
    monkeypatch.setenv("RFB_DFA_STATES", "0")
    assert "dfa 0 " in gpu_ctx.nfa_from_entries(snort.entries).describe()


def test_image_file_save_load_scan(gpu_ctx, snort, tmp_path):
    """an NFA saved as an execution-image file and loaded back (verified, not rebuilt) scans
    bit-identically; a multi-part NFA travels with its part table; a host-built file loads too."""
    data = WL.make_batch_numpy("wmix", snort.lo, snort.hi, 96, 1500, 1536, seed=0x5EED0077)
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    want = nfa.scan(data, 96, n_steps=1500, stride=1536)
    p = tmp_path / "snort.rfbimg"
    nfa.save_image(p)
    again = gpu_ctx.load_image(p)
    assert again.info == nfa.info
    got = again.scan(data, 96, n_steps=1500, stride=1536)
    assert recs_tuple(got.records) == recs_tuple(want.records) and np.array_equal(got.counts, want.counts)
    q = tmp_path / "host.rfbimg"
    R.image_file_build(snort.entries, q)
    assert q.read_bytes() == p.read_bytes()                       # the file is a function of the NFA alone
    E3, n3 = WL.replicate_nfa(snort.entries, snort.n_states, 3)
    big = gpu_ctx.nfa_from_entries(E3, n3)
    assert big.info["n_parts"] >= 2
    r3 = tmp_path / "x3.rfbimg"
    big.save_image(r3)
    big2 = gpu_ctx.load_image(r3)
    assert big2.info == big.info
    a = big.scan(data[:32], 32, n_steps=1500, stride=1536)
    b = big2.scan(data[:32], 32, n_steps=1500, stride=1536)
    assert recs_tuple(a.records) == recs_tuple(b.records) and a.n_matches == 3 * nfa.scan(data[:32], 32, n_steps=1500, stride=1536).n_matches
    with pytest.raises(R.RfbError):
        gpu_ctx.load_image(tmp_path / "missing.rfbimg")


@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_calibration_changes_no_result(gpu_ctx, snort, l7, name):
    """rfb_nfa_calibrate renumbers the start-DFA states by measured visit frequency (device tables only): records,
    counts and exported final sets must be what they were before, and still equal the oracle's."""
    rs = snort if name == "snort_16" else l7
    nfa = gpu_ctx.nfa_from_entries(rs.entries)
    n, L = 1024, 777
    data = WL.make_batch_numpy("wmix", rs.lo, rs.hi, n, L, 800, seed=0x5EED0300)
    assert nfa.calibration()[0] is False
    before = nfa.scan(data, n, n_steps=L, stride=800, want_state=True, state_cap=127)
    sample = WL.make_batch_numpy("whi", rs.lo, rs.hi, 256, 1500, 1536, seed=0x5EED0301)
    nfa.calibrate(sample, 256, 1500, 1536)
    done, symbols, hot = nfa.calibration()
    assert done and symbols == 256 * 1499 and 0.0 < hot <= 1.0
    after = nfa.scan(data, n, n_steps=L, stride=800, want_state=True, state_cap=127)
    assert recs_tuple(after.records) == recs_tuple(before.records) and np.array_equal(after.counts, before.counts)
    for s in range(n):
        assert sorted(after.state[s, 1: 1 + after.state[s, 0]].tolist()) == sorted(before.state[s, 1: 1 + before.state[s, 0]].tolist())
    want = O.b_scan_many(rs.entries, rs.n_states, data, n, 800, L)
    assert recs_tuple(after.records) == recs_tuple(want["recs"]) and np.array_equal(after.counts, want["counts"])
    # calibrating again on other traffic is allowed at any time
    nfa.calibrate(WL.make_batch_numpy("uniform", None, None, 64, 1500, 1536), 64, 1500, 1536)
    again = nfa.scan(data, n, n_steps=L, stride=800)
    assert recs_tuple(again.records) == recs_tuple(want["recs"])


def test_auto_calibration_on_the_first_large_batch(gpu_ctx, snort):
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    n = 8192
    data = WL.make_batch_numpy("wmix", snort.lo, snort.hi, n, 300, 320, seed=0x5EED0302)
    got = nfa.scan(data, n, n_steps=300, stride=320)
    assert nfa.calibration()[0] is True
    want = O.b_scan_many(snort.entries, snort.n_states, data, n, 320, 300)
    assert recs_tuple(got.records) == recs_tuple(want["recs"]) and np.array_equal(got.counts, want["counts"])


def test_auto_calibration_sample_does_not_alias_with_the_batch(gpu_ctx, snort):
    """W-mix alternates quiet (even) and busy (odd) streams.  A sample taken at a fixed even stride saw only the quiet
    half (hot fraction 0.998, the busy half ran on cold rows); the hashed sample must see both kinds: its hot fraction lies
    between the busy-only and the quiet-only one.  Host and device batches are sampled alike."""
    import torch
    n, L, stride = 16384, 1500, 1536
    frac = {}
    for mix in ("wlo", "whi", "wmix"):
        nfa = gpu_ctx.nfa_from_entries(snort.entries)
        data = WL.make_batch_numpy(mix, snort.lo, snort.hi, n, L, stride, seed=0x5EED0303)
        nfa.scan(data, n, n_steps=L, stride=stride, record_capacity=0)
        done, symbols, frac[mix] = nfa.calibration()
        assert done and symbols == 2048 * (L - 1)
    assert frac["whi"] + 0.02 < frac["wmix"] < frac["wlo"] - 0.02, frac
    nfa = gpu_ctx.nfa_from_entries(snort.entries)
    dev = torch.from_numpy(data).to("cuda:0")
    counts = torch.zeros(snort.n_states, dtype=torch.int64, device="cuda:0")
    torch.cuda.synchronize()
    nfa.scan_device(dev.data_ptr(), dev.numel(), n, L, stride, counts.data_ptr(), None, 0, flags=0)
    torch.cuda.synchronize()
    assert nfa.calibration() == (True, 2048 * (L - 1), frac["wmix"])


def test_two_shards_through_two_contexts_equal_one_shot(snort):
    """BASELINE config 4 on one GPU: the batch is cut into two contiguous shards, each scanned by its OWN context (its own
    copy of the NFA) with stream_id_base; merged counts and records must equal the one-shot scan and the oracle
    (streams are independent: Design/FPGA.v:54-57,264-268)."""
    from regex_fpga_b200 import shard
    n, L, stride = 3001, 500, 512
    data = WL.make_batch_numpy("wmix", snort.lo, snort.hi, n, L, stride, seed=0x5EED0400)
    want = O.b_scan_many(snort.entries, snort.n_states, data, n, stride, L)
    with R.Context(0) as one:
        whole = one.nfa_from_entries(snort.entries).scan(data, n, n_steps=L, stride=stride)
    assert recs_tuple(whole.records) == recs_tuple(want["recs"]) and np.array_equal(whole.counts, want["counts"])
    counts, recs = np.zeros(snort.n_states, np.uint64), []
    ctxs = [R.Context(0), R.Context(0)]
    try:
        for r, ctx in enumerate(ctxs):
            first, cnt = shard.shard_range(n, r, 2)
            nfa = ctx.nfa_from_entries(snort.entries)
            part = nfa.scan(data[first:first + cnt], cnt, n_steps=L, stride=stride, stream_id_base=first)
            counts += part.counts
            recs += recs_tuple(part.records)
    finally:
        for ctx in ctxs:
            ctx.close()
    assert np.array_equal(counts, want["counts"])
    assert recs == recs_tuple(want["recs"])          # shards are ascending and contiguous: concatenation is canonical order


def test_group_scan_in_one_process(snort):
    """rfb_group_*: N GPUs in one process: contiguous shards, NCCL all-reduce of the counts, records in shard
    order.  One GPU always (a 1-rank communicator still goes through ncclAllReduce); every visible GPU when there are more."""
    import torch
    n, L, stride = 10001, 400, 416
    data = WL.make_batch_numpy("wmix", snort.lo, snort.hi, n, L, stride, seed=0x5EED0500)
    want = O.b_scan_many(snort.entries, snort.n_states, data, n, stride, L)
    sizes = [1] + ([torch.cuda.device_count()] if torch.cuda.device_count() > 1 else [])
    for g in sizes:
        with R.Group(list(range(g))) as grp:
            assert grp.size == g
            nfa = grp.nfa_from_entries(snort.entries)
            for _ in range(2):                                           # the second call reuses every staging buffer
                got = nfa.scan(data, n, L, stride, record_capacity=1 << 18, stream_id_base=7)
                assert got.n_matches == want["n_recs"] and got.n_dropped == 0 and got.n_symbols == n * L
                assert np.array_equal(got.counts, want["counts"])
                w = want["recs"]
                assert recs_tuple(got.records) == list(zip((w["stream"] + 7).tolist(), w["pos"].tolist(), w["state"].tolist()))
            small = nfa.scan(data, n, L, stride, record_capacity=100)   # overflow is counted, never undefined
            assert small.n_records == 100 and small.n_dropped == want["n_recs"] - 100 and np.array_equal(small.counts, want["counts"])
            nfa.close()
