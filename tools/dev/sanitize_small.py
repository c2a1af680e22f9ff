"""Small scan for compute-sanitizer (memcheck / racecheck): both kernels, ragged + unaligned streams, hand-over."""
import sys, os
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from nfa_gen import build_entries

z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
E, n, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
with R.Context(0) as ctx:
    nfa = ctx.nfa_from_entries(E)
    data = WL.make_batch_numpy("wmix", lo, hi, 96, 300, 301)           # odd stride: every alignment occurs
    for flags in (R.SCAN_SORT_RECORDS, R.SCAN_SORT_RECORDS | R.SCAN_FORCE_WARP):
        r = nfa.scan(data, 96, n_steps=300, stride=301, flags=flags)
        print("matches", r.n_matches)
    steps = (np.arange(96) * 3 % 301).astype(np.uint32)
    r = nfa.scan(data, 96, stride=301, steps=steps)
    print("ragged", r.n_matches, r.n_symbols)
    rows = [[(1, s) for s in range(1, 79)]] + [[(1, s + 1 if s + 1 < 79 else 1), (2, 79)] for s in range(1, 79)] + [[]]
    E2, n2 = build_entries(rows)
    nfa2 = ctx.nfa_from_entries(E2, n2)
    d2 = np.ones((40, 64), np.uint8); d2[:, ::5] = 2
    r = nfa2.scan(d2, 40, n_steps=64, stride=64)
    print("handover", r.n_matches, r.n_rescanned)
    print("cycles", nfa.fpga_cycles(lo, hi, 300))
