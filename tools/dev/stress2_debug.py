"""Dev: re-run one case of stress_parity2.py and show how the result differs from the oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams
def tup(r): return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))
seed0, i = int(sys.argv[1]), int(sys.argv[2])
rng = np.random.default_rng(700000 + seed0 + i)
os.environ["RFB_DFA_STATES"] = str(rng.choice([0, 40, 16384]))
(E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 300)), alphabet=int(rng.integers(2, 16)),
                          p_sticky=float(rng.choice([0.0, 0.1, 0.3])), max_fanout=int(rng.integers(1, 4)), unanchored=bool(rng.integers(0, 2)))
n1 = n
copies = int(np.ceil(40000 / (n - 1)))
E, n = WL.replicate_nfa(E, n, copies)
print("base states", n1, "copies", copies, "state-0 degree", int(E[1] - E[0]))
with R.Context(0) as ctx:
    nfa = ctx.nfa_from_entries(E, n)
    print(nfa.describe())
    ns = int(rng.integers(1, 60)); Lmax = int(rng.integers(1, 200))
    steps = rng.integers(0, Lmax + 1, size=ns).astype(np.uint32)
    gaps = rng.integers(0, 40, size=ns)
    offsets = np.zeros(ns, np.uint64); pos = 0
    for s in range(ns):
        pos += int(gaps[s]); offsets[s] = pos; pos += int(steps[s])
    data = random_streams(rng, syms, 1, pos + 16, p_alpha=float(rng.choice([0.7, 0.95])))[0]
    want = [O.b_scan(E, n, data[int(offsets[s]):int(offsets[s]) + int(steps[s])], int(steps[s]), stream_id=s, cap=1 << 18) for s in range(ns)]
    wrec = sorted(t for w in want for t in tup(w["recs"]))
    for name, flags in (("lane", R.SCAN_SORT_RECORDS), ("warp", R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP), ("lane-unsorted", 0)):
        for rep in range(3):
            got = nfa.scan(data, ns, stride=0, offsets=offsets, steps=steps, record_capacity=1 << 20, flags=flags)
            g = sorted(tup(got.records))
            missing = sorted(set(wrec) - set(g)); extra = sorted(set(g) - set(wrec))
            print(name, rep, "n", len(g), "want", len(wrec), "dups", len(g) - len(set(g)), "missing", missing[:5], "extra", extra[:5], "n_symbols", got.n_symbols, int(steps.sum()), "rescanned", got.n_rescanned)
