"""Dev check: calibrate on one kind of traffic, scan another, compare with oracle B (single snort_16 and 7x)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
E, n, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
copies = int(sys.argv[2]) if len(sys.argv) > 2 else 1
EE, nn = (WL.replicate_nfa(E, n, copies) if copies > 1 else (E, n))
adv = WL.make_adversarial_numpy(E, n, hi, 2048, 1500, 1536)
data = WL.make_batch_numpy("whi", lo, hi, N, 1500, 1536, seed=0x5EED0005)
want = O.b_scan_many(EE, nn, data, N, 1536, 1500, want_recs=False)
ctx = R.Context(0); nfa = ctx.nfa_from_entries(EE, nn)
for label, sample in (("uncalibrated", None), ("calibrated on adv", adv), ("calibrated on whi", data[:2048])):
    if sample is not None:
        nfa.calibrate(sample, 2048, 1500, 1536)
    got = nfa.scan(data, N, n_steps=1500, stride=1536, record_capacity=0, flags=0)
    print(label, nfa.calibration(), "matches", got.n_matches, "oracle", want["n_recs"], "states differing", int(np.count_nonzero(got.counts != want["counts"])))
