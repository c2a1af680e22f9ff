"""Pins the CPU oracle (oracle/) -- the checker every GPU parity test relies on.

The reference ships no golden vectors, so the pins are: (1) the known-answer digests of
SURVEY.md Appendix C, reproduced here from the shipped data files; (2) the cycle-level restatement of
Design/FPGA.v + testbench (oracle A) agreeing with the functional restatement (oracle B) state by state;
(3) the closed-form cycle model agreeing with oracle A's clock.
"""
import hashlib

import numpy as np
import pytest

from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams

# SURVEY.md Appendix C (first 32 hex chars of SHA-256 over the canonical text form)
APPENDIX_C = {
    "snort_16": {"lo": ("ec9cf80a3c03258850f8d7a6bf755115", "ef2bdc9f522758f9b33dfcc65e5a9a27", 13),
                 "hi": ("4755bd6212512a38e35629c4ca6a9da6", "204f27f7bb99a691dabe57777ec4311d", 678),
                 "cycles": 2188184738},
    "l7_filter": {"lo": ("5c01e93bca12a25f71e41b58d30a997c", "23787b963669ec25976c84bb434f2650", 7),
                  "hi": ("f78ab982ae2433c413755f2cd91751cd", "491baa45d2b9b21a00b6bff3e5b51b2c", 5),
                  "cycles": 617518104},
}


def digest_counts(counts):
    s = "".join(f"{i} {int(counts[i])}\n" for i in np.nonzero(counts)[0])
    return hashlib.sha256(s.encode()).hexdigest()[:32]


def digest_events(recs):
    s = "".join(f"{int(p)} {int(st)}\n" for p, st in zip(recs["pos"], recs["state"]))
    return hashlib.sha256(s.encode()).hexdigest()[:32]


@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_oracle_b_reproduces_appendix_c(name, snort, l7, expected):
    rs = snort if name == "snort_16" else l7
    M = expected["tb_trace_entries"]
    for key, data in (("lo", rs.lo), ("hi", rs.hi)):
        b = O.b_scan(rs.entries, rs.n_states, data, M - 1)
        cd, ed, n = APPENDIX_C[name][key]
        assert b["n_recs"] == n
        assert digest_counts(b["counts"]) == cd
        assert digest_events(b["recs"]) == ed
        exp = expected["rulesets"][name]["tb"][key]
        assert exp["counts_digest"] == cd and exp["events_digest"] == ed
        assert [[int(p), int(s)] for p, s in zip(b["recs"]["pos"], b["recs"]["state"])] == exp["events"]


def test_appendix_c_spot_values(snort, l7):
    b = O.b_scan(snort.entries, snort.n_states, snort.lo, 199999)
    ev = list(zip(b["recs"]["pos"].tolist(), b["recs"]["state"].tolist()))
    assert ev[:4] == [(2569, 346), (3446, 200), (7759, 1499), (7975, 205)] and ev[-1] == (190402, 955)
    assert b["max_active"] == 10 and abs(b["sum_active"] / 199999 - 1.77) < 0.01
    b = O.b_scan(snort.entries, snort.n_states, snort.hi, 199999)
    assert b["counts"][200] == 20 and b["max_active"] == 37
    b = O.b_scan(l7.entries, l7.n_states, l7.hi, 199999)
    assert list(zip(b["recs"]["pos"].tolist(), b["recs"]["state"].tolist())) == \
        [(30, 2575), (1021, 443), (1832, 443), (5595, 443), (9672, 1386)]
    # full l7 files (262144 entries): one extra lo match on state 1109
    b = O.b_scan(l7.entries, l7.n_states, l7.lo, 262143)
    assert b["n_recs"] == 8 and b["counts"][1109] == 2


@pytest.mark.parametrize("name", ["snort_16", "l7_filter"])
def test_oracle_a_matches_b_and_cycle_model(name, snort, l7, expected):
    """Every-edge simulation of FPGA.v on a 2000-entry prefix: same pulses as the functional oracle,
    10-bit counters as TB:21-22, cycle total equal to the closed form and to the golden value."""
    rs = snort if name == "snort_16" else l7
    M = 2000
    a = O.a_run(rs.entries, rs.n_states, rs.lo, rs.hi, M, fast_idle=False)
    for stream, data, cnt, mc in ((0, rs.lo, a["counts1"], a["mc1"]), (1, rs.hi, a["counts2"], a["mc2"])):
        b = O.b_scan(rs.entries, rs.n_states, data, M - 1)
        assert np.array_equal(cnt, b["counts"])
        assert np.array_equal(mc, (b["counts"] & 0x3FF).astype(np.uint16))
        ev = a["recs"][a["recs"]["stream"] == stream]
        assert np.array_equal(ev["pos"], b["recs"]["pos"]) and np.array_equal(ev["state"], b["recs"]["state"])
    assert a["cycles"] == expected["rulesets"][name]["tb"]["cycles_first_2000_entries"]
    assert a["cycles"] == O.cycle_model(rs.entries, rs.n_states, rs.lo, rs.hi, M)
    fast = O.a_run(rs.entries, rs.n_states, rs.lo, rs.hi, M, fast_idle=True)
    assert fast["cycles"] == a["cycles"] and np.array_equal(fast["counts2"], a["counts2"])
    assert np.array_equal(fast["recs"], a["recs"])


def test_oracle_a_full_testbench_l7(l7, expected):
    """The whole committed testbench configuration (size_range = 2794, M = 200000; TB:20,71)."""
    a = O.a_run(l7.entries, l7.n_states, l7.lo, l7.hi, 200000, fast_idle=True)
    assert a["cycles"] == APPENDIX_C["l7_filter"]["cycles"] == expected["rulesets"]["l7_filter"]["tb"]["total_cycles"]
    assert digest_counts(a["counts1"]) == APPENDIX_C["l7_filter"]["lo"][0]
    assert digest_counts(a["counts2"]) == APPENDIX_C["l7_filter"]["hi"][0]


def test_expected_json_snort_cycles(expected):
    assert expected["rulesets"]["snort_16"]["tb"]["total_cycles"] == APPENDIX_C["snort_16"]["cycles"]


@pytest.mark.parametrize("seed", range(12))
def test_oracle_a_equals_b_on_random_nfas(seed):
    """The FSM's 3-deep line pipeline applies every CSR entry of an active row exactly once for any row
    length and alignment: random NFAs with rows of 0..300 entries."""
    rng = np.random.default_rng(1000 + seed)
    (E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 60)), alphabet=int(rng.integers(2, 12)))
    data = random_streams(rng, syms, 2, 400)
    M = 400
    a = O.a_run(E, n, data[0], data[1], M, fast_idle=bool(seed & 1))
    for stream in (0, 1):
        b = O.b_scan(E, n, data[stream], M - 1)
        cnt = a["counts1"] if stream == 0 else a["counts2"]
        assert np.array_equal(cnt, b["counts"])
        ev = a["recs"][a["recs"]["stream"] == stream]
        assert np.array_equal(ev["pos"], b["recs"]["pos"]) and np.array_equal(ev["state"], b["recs"]["state"])
    assert a["cycles"] == O.cycle_model(E, n, data[0], data[1], M)


def test_oracle_edge_cases():
    from nfa_gen import build_entries
    # single accepting start state: pulses on every step (size == 1 keeps input_char_flag high)
    E, n = build_entries([[]])
    b = O.b_scan(E, n, np.zeros(10, np.uint8), 5)
    assert b["n_recs"] == 1 and b["recs"]["pos"][0] == 0     # S_1 is empty: {0} has no successors
    a = O.a_run(E, n, np.zeros(10, np.uint8), np.zeros(10, np.uint8), 6)
    assert a["counts1"][0] == 1 and a["counts2"][0] == 1
    # zero steps
    E, n = build_entries([[(1, 1)], []])
    assert O.b_scan(E, n, np.zeros(4, np.uint8), 0)["n_recs"] == 0
    # 10-bit counter wrap (TB:21-22): state 1 accepts, 0 -> {0,1} on symbol 7 for ever
    E, n = build_entries([[(7, 0), (7, 1)], []])
    M = 1100
    d = np.full(M, 7, np.uint8)
    a = O.a_run(E, n, d, d, M)
    assert a["counts1"][1] == M - 2 and a["mc1"][1] == (M - 2) & 0x3FF
    many = O.b_scan_many(E, n, np.stack([d, d, d]), 3, M, M - 1, n_threads=2)
    assert many["counts"][1] == 3 * (M - 2) and many["n_recs"] == 3 * (M - 2)
    assert np.array_equal(many["recs"]["stream"], np.repeat(np.arange(3, dtype=np.uint32), M - 2))
