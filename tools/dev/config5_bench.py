"""BASELINE config 5 timing: 7 x snort_16 (66 592 states) on the general kernel, adversarial + hi-window streams."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as WL

z = np.load(os.path.join(ROOT, "tests", "golden", "snort_16.npz"))
E, n, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
E7, n7 = WL.replicate_nfa(E, n, 7)
ctx = R.Context(0)
t0 = time.time(); nfa = ctx.nfa_from_entries(E7, n7); print("load+verify s", round(time.time() - t0, 2), nfa.info)
N = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
for mix in ("adv", "whi"):
    if mix == "adv":
        batch = WL.make_adversarial_torch(E, n, hi, N, "cuda:0")
    else:
        batch = WL.make_batch_torch("whi", lo, hi, N, "cuda:0", seed=0x5EED0005)
    counts = torch.zeros(n7, dtype=torch.int64, device="cuda:0")
    torch.cuda.synchronize()   # the generator ran on the default stream
    st = torch.cuda.Stream(); torch.cuda.set_stream(st)
    for _ in range(3):
        r = nfa.scan_device(batch.data_ptr(), batch.numel(), N, 1500, 1536, counts.data_ptr(), None, 0, cuda_stream=st.cuda_stream)
    print(mix, "streams", N, "gpu_ms", round(r.gpu_ms, 2), "Gbit/s", round(N * 1500 * 8 / r.gpu_ms / 1e6, 2), "matches", r.n_matches)
