"""Dev tool: loop statistics of the lane kernel (needs the RFB_STATS variant: tools/dev/build_variant.sh stats -DRFB_STATS,
run with RFB_LIB=.../librfb200_stats.so).  The variant adds its counters into the last seven entries of the count vector."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
mix = sys.argv[1] if len(sys.argv) > 1 else "whi"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 131072
z = np.load("tests/golden/snort_16.npz")
E, ns, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
ctx = R.Context(0); nfa = ctx.nfa_from_entries(E)
batch = WL.make_adversarial_torch(E, ns, hi, n, "cuda:0", 1500, 1536) if mix == "adv" else WL.make_batch_torch(mix, lo, hi, n, "cuda:0", 1500, 1536)
counts = torch.zeros(ns, dtype=torch.int64, device="cuda:0")
torch.cuda.synchronize()
for it in range(2):
    counts.zero_()
    r = nfa.scan_device(batch.data_ptr(), batch.numel(), n, 1500, 1536, counts.data_ptr(), None, 0, flags=R.SCAN_ACCUMULATE)
    torch.cuda.synchronize()
c = counts[-7:].cpu().numpy().astype(np.float64)
sym = n * 1500.0
names = ["lane-iterations", "quiet runs", "quiet steps", "quiet runs of 0 steps", "general steps", "drain items", "quiet runs with a full chunk"]
print(mix, "streams", n, "gpu_ms", r.gpu_ms)
for nm, v in zip(names, c):
    print(f"  {nm:32s} {v:14.0f}  per stream {v / n:9.1f}  per symbol {v / sym:.4f}")
print("  calibration (measured, sample symbols, hot fraction):", nfa.calibration())
