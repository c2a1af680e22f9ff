"""Break down the host-pointer path (rfb_scan) on the bench workload: wall time per call for a few variants."""
import sys, time, numpy as np, torch
sys.path.insert(0, ".")
import regex_fpga_b200 as R
from regex_fpga_b200 import workloads as W
from regex_fpga_b200.engine import MATCH_DTYPE
g = np.load("tests/golden/snort_16.npz")
n = 1048576
with R.Context(0) as ctx:
    nfa = ctx.nfa_from_entries(g["entries"], int(g["n_states"]))
    batch = W.make_batch_torch("wmix", g["lo"], g["hi"], n, "cuda")
    host = torch.empty((n, 1536), dtype=torch.uint8, pin_memory=True); host.copy_(batch); torch.cuda.synchronize()
    hn = host.numpy()
    cap = 4 << 20
    prec = torch.empty(cap * 12, dtype=torch.uint8, pin_memory=True).numpy().view(MATCH_DTYPE)
    pcnt = torch.empty(nfa.n_states, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    def run(label, **kw):
        nfa.scan(hn, n, n_steps=1500, stride=1536, **kw)
        ts = []
        for _ in range(3):
            t0 = time.perf_counter(); out = nfa.scan(hn, n, n_steps=1500, stride=1536, **kw); ts.append(time.perf_counter() - t0)
        print(f"{label:40s} wall {min(ts)*1e3:7.2f} ms  gpu_ms {out.gpu_ms:7.2f}  recs {out.n_records}", flush=True)
    run("sorted, pageable results", record_capacity=cap, flags=1)
    run("sorted, pinned results", records_out=prec, counts_out=pcnt, flags=1)
    run("unsorted, pinned results", records_out=prec, counts_out=pcnt, flags=0)
    import os
    for ch, mb in ((16, 64), (32, 32), (64, 16), (128, 8), (256, 4)):
        os.environ["RFB_CHUNKS"] = str(ch); os.environ["RFB_CHUNK_MB"] = str(mb)
        run(f"sorted, pinned, chunks={ch}", records_out=prec, counts_out=pcnt, flags=1)
    del os.environ["RFB_CHUNKS"], os.environ["RFB_CHUNK_MB"]
    # pipelined: two batches in flight
    prec2 = torch.empty(cap * 12, dtype=torch.uint8, pin_memory=True).numpy().view(MATCH_DTYPE)
    pcnt2 = torch.empty(nfa.n_states, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64)
    bufs = [(prec, pcnt), (prec2, pcnt2)]
    for flags in (1, 0):
        K = 6
        t0 = time.perf_counter()
        for i in range(K):
            nfa.submit(hn, n, 1500, 1536, bufs[i & 1][0], bufs[i & 1][1], flags=flags)
            if i: out = nfa.wait()
        out = nfa.wait()
        dt = (time.perf_counter() - t0) / K
        print(f"pipelined (flags={flags})                    wall {dt*1e3:7.2f} ms per batch  = {n*1500*8/dt/1e9:6.1f} Gbit/s  gpu_ms {out.gpu_ms:7.2f} recs {out.n_records}", flush=True)
    run("no records, pinned counts", record_capacity=0, counts_out=pcnt, flags=0)
    run("no records, no counts", record_capacity=0, want_counts=False, flags=4)
