"""Second parity soak (dev tool): ragged batches (offsets + per-stream lengths incl. empty streams), both kernels,
resumed scans through the general kernel, NFAs cut into parts, unsorted records, small record buffers.
python tools/dev/stress_parity2.py [n_cases] [first_seed]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
from nfa_gen import random_nfa, random_streams

def tup(r): return list(zip(r["stream"].tolist(), r["pos"].tolist(), r["state"].tolist()))

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 100
seed0 = int(sys.argv[2]) if len(sys.argv) > 2 else 0
bad = 0
def report(i, what, nfa):
    global bad
    bad += 1
    print("MISMATCH case", i, what, nfa.describe().strip().replace("\n", " | ")[:300], flush=True)

with R.Context(0) as ctx:
    for i in range(n_cases):
        rng = np.random.default_rng(700000 + seed0 + i)
        os.environ["RFB_DFA_STATES"] = str(rng.choice([0, 40, 16384]))
        (E, n), syms = random_nfa(rng, n_states=int(rng.integers(3, 300)), alphabet=int(rng.integers(2, 16)),
                                  p_sticky=float(rng.choice([0.0, 0.1, 0.3])), max_fanout=int(rng.integers(1, 4)),
                                  unanchored=bool(rng.integers(0, 2)))
        big = i % 5 == 4 and n >= 40 and not os.environ.get("STRESS_NO_BIG")
        if big:   # enough replicas that the tables no longer fit one SM: the NFA is cut into parts
            copies = int(np.ceil(40000 / (n - 1)))
            try:
                E, n = WL.replicate_nfa(E, n, copies)
            except AssertionError:
                big = False
        nfa = ctx.nfa_from_entries(E, n)
        ns = int(rng.integers(1, 60)); Lmax = int(rng.integers(1, 200))
        steps = rng.integers(0, Lmax + 1, size=ns).astype(np.uint32)
        gaps = rng.integers(0, 40, size=ns)
        offsets = np.zeros(ns, np.uint64); pos = 0
        for s in range(ns):
            pos += int(gaps[s]); offsets[s] = pos; pos += int(steps[s])
        data = random_streams(rng, syms, 1, pos + 16, p_alpha=float(rng.choice([0.7, 0.95])))[0]
        want = [O.b_scan(E, n, data[int(offsets[s]):int(offsets[s]) + int(steps[s])], int(steps[s]), stream_id=s, cap=1 << 20) for s in range(ns)]
        if any(w["n_recs"] > (1 << 20) for w in want) or sum(w["n_recs"] for w in want) > (1 << 22):
            # the checker's own buffers would truncate (a replicated NFA multiplies every match by its copies): not a case
            print("SKIP case", i, "more records than the checker keeps", flush=True)
            continue
        wrec = sorted(t for w in want for t in tup(w["recs"]))
        wcnt = sum(w["counts"] for w in want)
        for name, flags in (("lane", R.SCAN_SORT_RECORDS), ("warp", R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP), ("lane-unsorted", 0)):
            got = nfa.scan(data, ns, stride=0, offsets=offsets, steps=steps, record_capacity=1 << 22, flags=flags)
            g = tup(got.records)
            if (sorted(g) if flags == 0 else g) != wrec or not np.array_equal(got.counts, wcnt) or got.n_symbols != int(steps.sum()):
                report(i, name, nfa)
        # small record buffer: the overflow is counted, the kept records are a subset, counts stay exact
        capr = max(1, len(wrec) // 3)
        got = nfa.scan(data, ns, stride=0, offsets=offsets, steps=steps, record_capacity=capr, flags=0)
        if got.n_matches != len(wrec) or got.n_dropped != len(wrec) - min(capr, len(wrec)) or not set(tup(got.records)) <= set(wrec) or not np.array_equal(got.counts, wcnt):
            report(i, "small-buffer", nfa)
        # resumed through either kernel on a uniform sub-batch
        L = int(rng.integers(2, 120)); m = int(rng.integers(1, 40))
        d2 = random_streams(rng, syms, m, L, p_alpha=0.9)
        w2 = O.b_scan_many(E, n, d2, m, L, L, cap=1 << 20)
        cut = L // 2
        for name, flags in (("lane", R.SCAN_SORT_RECORDS), ("warp", R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP)):
            a = nfa.scan(np.ascontiguousarray(d2[:, :cut]), m, n_steps=cut, stride=cut, want_state=True, state_cap=255, flags=flags)
            if np.any(a.state[:, 0] == R.STATE_OVERFLOW): continue
            b = nfa.scan(np.ascontiguousarray(d2[:, cut:]), m, n_steps=L - cut, stride=L - cut, state_in=a.state, pos_base=cut, flags=flags)
            if sorted(tup(a.records) + tup(b.records)) != tup(w2["recs"]):
                report(i, "resumed-" + name + ("-parts" if big else ""), nfa)
print("cases", n_cases, "mismatches", bad)
sys.exit(1 if bad else 0)
