"""Dev check: config 5 (7 x snort_16) counts of the one-launch path / the pass-per-part path against oracle B."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
from oracle import oracle_py as O
z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
E, n, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
E7, n7 = WL.replicate_nfa(E, n, 7)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
data = WL.make_batch_numpy("whi", lo, hi, N, 1500, 1536, seed=0x5EED0005)
want = O.b_scan_many(E7, n7, data, N, 1536, 1500, want_recs=False)
ctx = R.Context(0); nfa = ctx.nfa_from_entries(E7, n7)
import torch
dev_batch = WL.make_batch_torch("whi", lo, hi, N, "cuda:0", seed=0x5EED0005)
assert np.array_equal(dev_batch.cpu().numpy(), data), "torch and numpy generators disagree"
cnt = torch.zeros(n7, dtype=torch.int64, device="cuda:0")
for rep in range(3):
    r = nfa.scan_device(dev_batch.data_ptr(), dev_batch.numel(), N, 1500, 1536, cnt.data_ptr(), None, 0)
    dc = cnt.cpu().numpy().astype(np.uint64)
    print("device path rep", rep, "matches", r.n_matches, "oracle", want["n_recs"], "rescanned", r.n_rescanned, "states differing", int(np.count_nonzero(dc != want["counts"])))
for rep in range(3):
    got = nfa.scan(data, N, n_steps=1500, stride=1536, record_capacity=0, flags=0)
    bad = np.nonzero(got.counts != want["counts"])[0]
    print("multi" if not os.environ.get("RFB_NO_MULTI") else "per-part", "rep", rep, "matches", got.n_matches, "oracle", want["n_recs"], "rescanned", got.n_rescanned,
          "states differing", bad.size, bad[:8].tolist(), (got.counts[bad[:8]].astype(np.int64) - want["counts"][bad[:8]].astype(np.int64)).tolist())
