"""Dev measurement: the same W-mix streams in their normal order (lo/hi alternate, so every warp holds both kinds) and
sorted by kind (all lo windows first: warps are homogeneous most of the time).  The difference is the mixing penalty."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
E, ns, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
ctx = R.Context(0); nfa = ctx.nfa_from_entries(E)
batch = WL.make_batch_torch("wmix", lo, hi, n, "cuda:0", 1500, 1536).view(n, 1536)
idx = torch.arange(n, device="cuda:0")
orders = {"alternating": idx, "sorted by kind": torch.cat([idx[0::2], idx[1::2]]),
          "blocks of 32": idx.view(-1, 2, 32).transpose(1, 2).reshape(-1) if False else torch.cat([idx.view(-1, 64)[:, 0::2], idx.view(-1, 64)[:, 1::2]], dim=1).reshape(-1)}
counts = torch.zeros(ns, dtype=torch.int64, device="cuda:0")
for name, o in orders.items():
    b = batch[o].contiguous()
    torch.cuda.synchronize()
    ms = []
    for it in range(6):
        r = nfa.scan_device(b.data_ptr(), b.numel(), n, 1500, 1536, counts.data_ptr(), None, 0, flags=0)
        torch.cuda.synchronize()
        ms.append(r.gpu_ms)
    print(name, "ms", [round(x, 3) for x in ms[2:]], "matches", r.n_matches)
