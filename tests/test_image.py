"""Host logic of the load-time re-indexing (csrc/image.cpp): rfb_image_check builds the lane kernel's
execution image and proves it equivalent to the CSR for every (state, symbol) pair."""
import numpy as np
import pytest

import regex_fpga_b200 as R
from nfa_gen import build_entries, random_nfa


@pytest.mark.parametrize("sticky_words,bucket_bits", [(0, 0), (1, 3), (2, 4), (1, 2)])
def test_shipped_rulesets_verify(snort, l7, sticky_words, bucket_bits):
    for rs in (snort, l7):
        info = R.image_check(rs.entries, -1, sticky_words, bucket_bits)
        assert info["image_ok"] == 1
        assert info["n_states"] == rs.n_states
        assert info["image_bytes"] < 190 * 1024
        assert info["n_slots"] <= 0x8000


def test_shipped_ruleset_shapes(snort, l7, monkeypatch):
    s = R.image_check(snort.entries)
    assert (s["n_states"], s["n_transitions"], s["n_accepting"]) == (9514, 79856, 536)
    assert 50 <= s["n_sticky"] < 65 and s["sticky_words"] == 1   # some of the 67 moved into the start DFA: one mask word is enough
    monkeypatch.setenv("RFB_DFA_ABSORB", "0")
    s = R.image_check(snort.entries)
    assert s["n_sticky"] >= 65          # 23 full + 42 all-but-newline self-loop states
    monkeypatch.delenv("RFB_DFA_ABSORB")
    f = R.image_check(l7.entries)
    assert (f["n_states"], f["n_transitions"], f["n_accepting"]) == (2794, 124977, 204)


@pytest.mark.parametrize("seed", range(40))
def test_random_nfas_verify(seed):
    rng = np.random.default_rng(seed)
    (E, n), _ = random_nfa(rng, n_states=int(rng.integers(2, 300)), alphabet=int(rng.integers(2, 40)),
                           p_sticky=float(rng.choice([0.0, 0.1, 0.5])), max_fanout=int(rng.integers(1, 4)))
    for sw, bb in ((0, 0), (1, 1), (2, 5)):
        info = R.image_check(E, n, sw, bb)
        assert info["image_ok"] == 1 and info["n_states"] == n


def test_more_sticky_candidates_than_mask_bits():
    """> 128 self-looping states: the surplus must stay correct as ordinary (class-edge) states."""
    rows = [[(1, s) for s in range(1, 200)]]
    for s in range(1, 200):
        rows.append([(c, s) for c in range(256) if c != s % 7] + [(2, (s % 198) + 1)])
    E, n = build_entries(rows)
    info = R.image_check(E, n)
    assert info["image_ok"] == 1 and info["n_sticky"] == 128 and info["sticky_words"] == 2


def test_invalid_images_are_rejected():
    E, n = build_entries([[(1, 1)], []])
    bad = E.copy()
    bad[n + 1] = (1 << 24) | 5            # target >= n_states
    with pytest.raises(R.RfbError) as e:
        R.image_check(bad, n)
    assert e.value.code == -4
    bad = E.copy()
    bad[1] = 3
    bad[2] = 1                            # row_ptr decreasing
    with pytest.raises(R.RfbError):
        R.image_check(bad, n)
    with pytest.raises(R.RfbError):
        R.image_check(np.array([1, 2, 3, 4], np.uint32), -1)   # row_ptr[0] != 0: size not detectable


def test_too_large_for_shared_memory_falls_to_warp_kernel():
    """40000 branching states do not fit the 15-bit id space: image_ok = 0, NFA still valid."""
    n = 40000
    rows = [[(1, (s + 1) % n or 1), (2, (s + 7) % n or 1)] for s in range(n)]
    E, n = build_entries(rows)
    info = R.image_check(E, n)
    assert info["image_ok"] == 0 and info["n_states"] == n


def test_replicated_nfa_is_valid_but_too_large_for_the_lane_tables(snort):
    """BASELINE config 5 image: 7 x snort_16 = 66 592 states; valid CSR, size auto-detectable, lane tables
    not buildable (more than 32 768 slots) -> general kernel."""
    from regex_fpga_b200 import workloads as WL
    E7, n7 = WL.replicate_nfa(snort.entries, snort.n_states, 7)
    assert n7 == 1 + 7 * 9513 and R.coe_detect_size(E7) == n7
    info = R.image_check(E7, n7)
    assert info["image_ok"] == 0 and info["n_accepting"] == 7 * 536 and info["n_transitions"] == 558992
    E1, n1 = WL.replicate_nfa(snort.entries, snort.n_states, 1)
    assert n1 == snort.n_states and np.array_equal(E1, snort.entries)


def test_adversarial_prefixes_reach_long_lived_states(snort):
    from regex_fpga_b200 import workloads as WL
    pref = WL.adversarial_prefixes(snort.entries, snort.n_states)
    assert len(pref) >= 40 and all(1 <= p.size < 200 for p in pref)


def test_hypothesis_random_nfas_verify():
    """Property-based: any well-formed CSR image (arbitrary rows, self-loops, fan-out, accept states, a start state
    that may itself be sticky or accepting) yields an execution image that passes the exhaustive verifier, or is
    cleanly reported as general-kernel-only."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @st.composite
    def nfas(draw):
        n = draw(st.integers(1, 24))
        rows = []
        for s in range(n):
            kind = draw(st.integers(0, 5))
            r = []
            if kind == 0:
                pass                                                   # accepting
            elif kind == 1:                                            # sticky-ish: long self loop + exits
                lo = draw(st.integers(0, 200))
                r += [(c, s) for c in range(lo, min(256, lo + draw(st.integers(16, 256))))]
                for _ in range(draw(st.integers(0, 3))):
                    r.append((draw(st.integers(0, 255)), draw(st.integers(0, n - 1)) or min(1, n - 1)))
            else:
                for _ in range(draw(st.integers(1, 6))):
                    c = draw(st.integers(0, 255))
                    t = draw(st.integers(0, n - 1))
                    if t == 0:
                        t = min(1, n - 1)                              # nothing targets state 0 in the shipped images
                    r.append((c, t))
                    if draw(st.booleans()):
                        r.append((c ^ 0x20, t))
            rows.append([(c, t) for c, t in r if not (t == 0 and n > 1)])
        return rows

    @settings(max_examples=60, deadline=None, suppress_health_check=[HealthCheck.too_slow])
    @given(nfas(), st.sampled_from([(0, 0), (1, 2), (2, 4)]))
    def check(rows, opt):
        E, n = build_entries(rows)
        info = R.image_check(E, n, opt[0], opt[1])
        assert info["n_states"] == n and info["image_ok"] == 1

    check()


@pytest.mark.parametrize("budget", ["0", "16", "150", "16384"])
def test_start_dfa_verifies_at_every_budget(snort, l7, budget, monkeypatch):
    """image_build proves the start DFA against the CSR for every (DFA state, symbol) before it accepts the image
    (image_verify): complete DFAs, DFAs cut at a small budget (failure-link fallback rows) and no DFA at all."""
    monkeypatch.setenv("RFB_DFA_STATES", budget)
    for rs in (snort, l7):
        info = R.image_check(rs.entries)
        assert info["image_ok"] == 1
    rng = np.random.default_rng(31)
    for _ in range(12):
        (E, n), _ = random_nfa(rng, n_states=int(rng.integers(3, 200)), alphabet=int(rng.integers(2, 12)),
                               p_sticky=0.1, unanchored=True)
        assert R.image_check(E, n)["image_ok"] == 1


def test_start_dfa_build_is_bounded_when_subsets_explode():
    """600 states over a 4-letter alphabet with random back edges: the subset construction would run away; the
    builder stops creating DFA states at a work bound and resolves the rest through failure links."""
    import time
    rng = np.random.default_rng(3)
    n = 600
    rows = [[(c, 1) for c in range(256)],
            [(c, 1) for c in range(256)] + [(int(rng.integers(97, 101)), int(rng.integers(2, n))) for _ in range(40)]]
    for _ in range(2, n):
        rows.append([(int(rng.integers(97, 101)), int(rng.integers(2, n))) for _ in range(4)])
    E, nn = build_entries(rows)
    t0 = time.time()
    info = R.image_check(E, nn)
    assert info["image_ok"] == 1 and time.time() - t0 < 60
