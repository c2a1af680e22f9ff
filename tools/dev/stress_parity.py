"""One-off parity soak (dev tool): many random NFAs (anchored and unanchored, several start-DFA budgets, resumed
halves) against oracle B.  python tools/dev/stress_parity.py [n_cases] [first_seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import regex_fpga_b200 as R
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams

def tup(r): return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 200
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad = 0
with R.Context(0) as ctx:
    for i in range(n_cases):
        rng = np.random.default_rng(500000 + seed0 + i)
        os.environ["RFB_DFA_STATES"] = str(rng.choice([0, 12, 40, 300, 16384]))
        (E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 500)), alphabet=int(rng.integers(2, 20)),
                                  p_sticky=float(rng.choice([0.0, 0.05, 0.2, 0.4])), p_accept=float(rng.choice([0.05, 0.15, 0.3])),
                                  max_fanout=int(rng.integers(1, 5)), unanchored=bool(rng.integers(0, 4)))
        nfa = ctx.nfa_from_entries(E, n)
        L = int(rng.integers(2, 400)); ns = int(rng.integers(1, 150))
        data = random_streams(rng, syms, ns, L, p_alpha=float(rng.choice([0.6, 0.85, 0.97])))
        want = O.b_scan_many(E, n, data, ns, L, L, cap=1 << 22)
        got = nfa.scan(data, ns, n_steps=L, stride=L, record_capacity=1 << 22)
        ok = got.n_matches == want["n_recs"] and np.array_equal(got.counts, want["counts"]) and tup(got.records) == tup(want["recs"])
        cut = L // 2
        if ok and cut > 0:
            a = nfa.scan(np.ascontiguousarray(data[:, :cut]), ns, n_steps=cut, stride=cut, want_state=True, state_cap=255, record_capacity=1 << 22)
            if not np.any(a.state[:, 0] == R.STATE_OVERFLOW):
                b = nfa.scan(np.ascontiguousarray(data[:, cut:]), ns, n_steps=L - cut, stride=L - cut, state_in=a.state, pos_base=cut, record_capacity=1 << 22)
                ok = sorted(tup(a.records) + tup(b.records)) == tup(want["recs"])
        if not ok:
            bad += 1
            print("MISMATCH case", i, "seed", 500000 + seed0 + i, "budget", os.environ["RFB_DFA_STATES"], nfa.describe().strip(), flush=True)
        nfa.close() if hasattr(nfa, "close") else None
print("cases", n_cases, "mismatches", bad)
sys.exit(1 if bad else 0)
