"""The execution image interpreted on the HOST exactly as the lane kernel interprets it (tests/image_replay.cpp,
test infrastructure) against oracle B: sticky masks, hashed rows and the start DFA with its insertion lists and
failure-link rows, in motion, without a GPU."""
import os
import subprocess

import numpy as np
import pytest

import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "regex_fpga_b200", "csrc")


@pytest.fixture(scope="session")
def replay_bin(tmp_path_factory):
    out = tmp_path_factory.mktemp("replay") / "image_replay"
    srcs = [os.path.join(ROOT, "tests", "image_replay.cpp")] + [os.path.join(CSRC, f) for f in ("image.cpp", "nfa.cpp", "formats.cpp")]
    subprocess.run(["g++", "-O2", "-std=c++17", "-I", CSRC, *srcs, "-o", str(out)], check=True)
    return str(out)


def replay(replay_bin, tmp_path, E, data2d, n_steps, budget):
    coe, binf = tmp_path / "n.coe", tmp_path / "d.bin"
    R.coe_write(coe, E, style=0)
    data2d = np.ascontiguousarray(data2d, dtype=np.uint8)
    data2d.tofile(binf)
    p = subprocess.run([replay_bin, str(coe), str(binf), str(data2d.shape[0]), str(data2d.shape[1]), str(n_steps), str(budget)],
                       check=True, capture_output=True, text=True)
    return [tuple(map(int, line.split())) for line in p.stdout.splitlines()]


def oracle(E, n, data2d, n_steps):
    w = O.b_scan_many(E, n, data2d, data2d.shape[0], data2d.shape[1], n_steps, cap=1 << 20)
    r = w["recs"]
    return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))


@pytest.mark.parametrize("budget", [0, 16, 200, 16384])
def test_replay_random_unanchored_nfas(replay_bin, tmp_path, budget):
    rng = np.random.default_rng(4200 + budget)
    for _ in range(6):
        (E, n), syms = random_nfa(rng, n_states=int(rng.integers(10, 250)), alphabet=int(rng.integers(3, 12)),
                                  p_sticky=float(rng.choice([0.0, 0.1, 0.3])), max_fanout=int(rng.integers(1, 4)), unanchored=True)
        L = int(rng.integers(20, 200))
        data = random_streams(rng, syms, int(rng.integers(4, 40)), L)
        assert replay(replay_bin, tmp_path, E, data, L, budget) == oracle(E, n, data, L)


@pytest.mark.parametrize("budget", [300, 16384])
def test_replay_shipped_rulesets(replay_bin, tmp_path, snort, l7, budget):
    for rs in (snort, l7):
        data = WL.make_batch_numpy("wmix", rs.lo, rs.hi, 24, 1500, 1536, seed=0x5EED0200)
        assert replay(replay_bin, tmp_path, rs.entries, data, 1500, budget) == oracle(rs.entries, rs.n_states, data, 1500)
