"""Execution-image files (csrc/imagefile.cpp): write, read back, and refuse anything that is not
a faithful re-indexing of the CSR it carries.  Host-only; the GPU side is in test_gpu_parity.py."""
import struct

import numpy as np
import pytest

import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from nfa_gen import random_nfa


def fnv1a64(b):
    h = 0xcbf29ce484222325
    for x in b:
        h = ((h ^ x) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return h


def reseal(b):
    """Recompute the trailing checksum so that only the structural / semantic checks stand between a tampered
    file and the kernel."""
    body = bytes(b[:-8])
    return body + struct.pack("<Q", fnv1a64(body))


def test_round_trip_shipped_rulesets(snort, l7, tmp_path):
    for name, rs in (("snort", snort), ("l7", l7)):
        p = tmp_path / f"{name}.rfbimg"
        R.image_file_build(rs.entries, p)
        info = R.image_file_check(p)
        assert info == R.image_check(rs.entries)            # same tables as a fresh build
        raw = p.read_bytes()
        assert raw[:8] == b"RFBIMG\x00\x01" and struct.unpack_from("<II", raw, 8) == (2, 1)
        n_entries, n_states = struct.unpack_from("<QI", raw, 16)
        assert n_states == rs.n_states and n_entries == rs.entries.size
        assert np.array_equal(np.frombuffer(raw, np.uint32, n_entries, 32), rs.entries)   # the .coe contents, verbatim


def test_multi_part_plan_round_trip(snort, tmp_path):
    E3, n3 = WL.replicate_nfa(snort.entries, snort.n_states, 3)      # 28 540 states: cut into parts
    p = tmp_path / "x3.rfbimg"
    R.image_file_build(E3, p, n3)
    info = R.image_file_check(p)
    assert info["n_parts"] >= 2 and info["image_ok"] == 1 and info["n_states"] == n3


def test_corruption_is_refused(l7, tmp_path):
    snort = l7                                                          # the smaller ruleset: every accepted or refused mutant costs a full proof
    p = tmp_path / "s.rfbimg"
    R.image_file_build(snort.entries, p)
    raw = bytearray(p.read_bytes())
    q = tmp_path / "bad.rfbimg"
    # 1. any flipped bit fails the checksum
    for at in (3, 40, len(raw) // 2, len(raw) - 9):
        b = bytearray(raw); b[at] ^= 0x10
        q.write_bytes(bytes(b))
        with pytest.raises(R.RfbError):
            R.image_file_check(q)
    # 2. truncation, trailing bytes
    q.write_bytes(bytes(raw[:-100]))
    with pytest.raises(R.RfbError):
        R.image_file_check(q)
    q.write_bytes(reseal(raw + b"\0" * 8))
    with pytest.raises(R.RfbError):
        R.image_file_check(q)
    # 3. a resealed file whose tables were edited: caught by the structural checks or by the proof against the CSR
    n_entries = struct.unpack_from("<Q", raw, 16)[0]
    tables = 32 + 4 * n_entries                                         # first byte after the CSR
    rng = np.random.default_rng(5)
    refused = 0
    for _ in range(40):
        b = bytearray(raw)
        at = int(rng.integers(tables, len(raw) - 8))
        b[at] ^= 1 << int(rng.integers(0, 8))
        q.write_bytes(reseal(b))
        try:
            R.image_file_check(q)
        except R.RfbError:
            refused += 1
    # bits that no lookup can observe exist (padding, unreachable table slots); everything that changes behaviour
    # must be refused, and most single-bit edits do
    assert refused >= 30
    # 4. an edited CSR under unchanged tables is refused too
    b = bytearray(raw)
    at = 32 + 4 * (snort.n_states + 1 + 17)                             # transition 17: flip its symbol
    b[at + 3] ^= 0x01
    q.write_bytes(reseal(b))
    with pytest.raises(R.RfbError):
        R.image_file_check(q)


def test_random_nfas_round_trip(tmp_path):
    rng = np.random.default_rng(77)
    for k in range(10):
        (E, n), _ = random_nfa(rng, n_states=int(rng.integers(2, 150)), alphabet=int(rng.integers(2, 12)),
                               p_sticky=0.15, unanchored=bool(k & 1))
        p = tmp_path / f"r{k}.rfbimg"
        R.image_file_build(E, p, n)
        assert R.image_file_check(p) == R.image_check(E, n)


def test_aliased_state_ids_are_refused(l7, tmp_path):
    """A resealed file in which a second table slot names an existing state (orig_of_id[alias] = s) and an edge is
    re-pointed at the alias: the per-(state, symbol) proof maps ids back through orig_of_id and would still see the
    right successor, while the kernel would follow the alias's arbitrary row.  The structural check demands that
    orig_of_id be the exact inverse of id_of_orig."""
    p = tmp_path / "a.rfbimg"
    R.image_file_build(l7.entries, p)
    raw = bytearray(p.read_bytes())
    n_entries = struct.unpack_from("<Q", raw, 16)[0]
    at = 32 + 4 * n_entries
    n_group, image_ok, header_bytes = struct.unpack_from("<III", raw, at)
    assert n_group == 0 and image_ok == 1
    at += 12
    hdr = struct.unpack_from("<%dI" % (header_bytes // 4), raw, at)
    n_slots, off_tab = hdr[0], hdr[12]
    at += header_bytes
    n_blob = struct.unpack_from("<Q", raw, at)[0]
    blob_at = at + 8
    at = blob_at + n_blob
    n_orig = struct.unpack_from("<Q", raw, at)[0]
    orig_at = at + 8
    assert n_orig == n_slots
    orig = np.frombuffer(bytes(raw[orig_at:orig_at + 4 * n_orig]), np.uint32).copy()
    tab = np.frombuffer(bytes(raw[blob_at + off_tab:blob_at + off_tab + 4 * n_slots]), np.uint32).copy()
    free = [i for i in range(n_slots - 1, 0, -1) if orig[i] == 0xFFFFFFFF]
    assert free, "the image has no unused slot to alias"
    alias = free[0]
    # an ordinary single-target record whose target is a real state
    victim = next(i for i in range(n_slots) if (tab[i] & 0xFF) <= ((tab[i] >> 8) & 0xFF) and orig[(tab[i] >> 16) & 0x7FFF] != 0xFFFFFFFF
                  and orig[i] != 0xFFFFFFFF and ((tab[i] >> 16) & 0x7FFF) >= 128)
    target = (int(tab[victim]) >> 16) & 0x7FFF
    tab[victim] = (int(tab[victim]) & 0x8000FFFF) | (alias << 16)
    orig[alias] = orig[target]
    raw[blob_at + off_tab:blob_at + off_tab + 4 * n_slots] = tab.tobytes()
    raw[orig_at:orig_at + 4 * n_orig] = orig.tobytes()
    q = tmp_path / "alias.rfbimg"
    q.write_bytes(reseal(raw))
    with pytest.raises(R.RfbError) as e:
        R.image_file_check(q)
    assert "inverse" in str(e.value)
