"""Debug helper for tools/dev/stress_parity.py: re-run given case seeds and print where the scan differs."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import regex_fpga_b200 as R
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams
def tup(r): return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))
with R.Context(0) as ctx:
    for seed in map(int, sys.argv[1:]):
        rng = np.random.default_rng(seed)
        budget = str(rng.choice([0, 12, 40, 300, 16384]))
        (E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 500)), alphabet=int(rng.integers(2, 20)),
                                  p_sticky=float(rng.choice([0.0, 0.05, 0.2, 0.4])), p_accept=float(rng.choice([0.05, 0.15, 0.3])),
                                  max_fanout=int(rng.integers(1, 5)), unanchored=bool(rng.integers(0, 4)))
        L = int(rng.integers(2, 400)); ns = int(rng.integers(1, 150))
        data = random_streams(rng, syms, ns, L, p_alpha=float(rng.choice([0.6, 0.85, 0.97])))
        want = O.b_scan_many(E, n, data, ns, L, L, cap=1 << 22)
        for bud in (budget, "0", "16384"):
            os.environ["RFB_DFA_STATES"] = bud
            nfa = ctx.nfa_from_entries(E, n)
            for name, flags in (("lane", 1), ("warp", 3)):
                got = nfa.scan(data, ns, n_steps=L, stride=L, record_capacity=1 << 22, flags=flags)
                g, w = tup(got.records), tup(want["recs"])
                same = g == w
                print(f"seed {seed} budget {bud} {name}: n={n} L={L} ns={ns} matches {got.n_matches}/{want['n_recs']} rescanned {got.n_rescanned} dropped {got.n_dropped} same={same}")
                if not same:
                    sg, sw = set(g), set(w)
                    miss = sorted(sw - sg)[:5]; extra = sorted(sg - sw)[:5]
                    print("   missing", miss, "extra", extra, "dups", len(g) - len(sg))
            cut = L // 2
            a = nfa.scan(np.ascontiguousarray(data[:, :cut]), ns, n_steps=cut, stride=cut, want_state=True, state_cap=255, record_capacity=1 << 22)
            ovf = int(np.sum(a.state[:, 0] == R.STATE_OVERFLOW))
            b = nfa.scan(np.ascontiguousarray(data[:, cut:]), ns, n_steps=L - cut, stride=L - cut, state_in=a.state, pos_base=cut, record_capacity=1 << 22)
            r = sorted(tup(a.records) + tup(b.records))
            print(f"   resumed: overflow rows {ovf} same={r == tup(want['recs'])} rescanned {a.n_rescanned}+{b.n_rescanned}")
            if r != tup(want["recs"]) and not ovf:
                sg, sw = set(r), set(tup(want["recs"]))
                print("   missing", sorted(sw - sg)[:5], "extra", sorted(sg - sw)[:5], "dups", len(r) - len(sg))
            print("  ", nfa.describe().strip())
