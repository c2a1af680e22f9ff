#!/usr/bin/env python
"""bench.py -- headline benchmark: Gbit/s scanned on the snort_16 NFA (BASELINE.json).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--mix wmix|whi|wlo|uniform]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch: 1 Mi independent 1500-byte packet streams per GPU
(BASELINE.json configs[2]; configs[3] at N > 1 -- weak scaling, N Mi streams in total, contiguous shards,
one NCCL all-reduce of the per-state match counts per step).  The batch (1.6 GB per GPU) is far larger
than the 126 MB L2, so every step streams its input from HBM.

  value        : whole-job Gbit/s with the batch already resident in HBM (CUDA events, max over ranks)
  e2e          : the same metric through the host-pointer API with HOST buffers (pinned H2D of the batch, kernels,
                 record sort, D2H of counts + match records inside the timed region), two batches in flight
                 (rfb_scan_submit / rfb_scan_wait)
  roofline     : HBM roofline of the scan kernel -- algorithmic bytes = 1 byte per symbol scanned
  cpu_baseline : the cycle-level CPU restatement of the reference design (oracle/oracle_a.c, the stand-in
                 for the "Verilated reference": no HDL simulator exists in this image) on a bounded sample
  --impl reference : that same CPU restatement as the timed arm, all host threads
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

STREAM_LEN = 1500
STRIDE = 1536          # device-resident batches: every stream starts 128-byte aligned
E2E_STRIDE = 1500      # host batches of the e2e leg: packed
SEED = 0x5EED0001
E2E_FLAGS = 1   # RFB_SCAN_SORT_RECORDS: the end-to-end call returns records in canonical (stream, pos, state) order
METRIC = "gbit_per_s_scanned_snort16"   # BASELINE.json: Gbit/s scanned (snort_16 NFA)


def metric_name(args):
    return METRIC if args.ruleset == "snort_16" else "gbit_per_s_scanned_" + args.ruleset


def load_ruleset(name="snort_16"):
    z = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return z["entries"], int(z["n_states"]), z["lo"], z["hi"]


def ncu_capture():
    """Per-symbol figures of the lane kernel from the committed ncu capture of the shipped build on this workload
    (profiles/r2_kernel.json: DRAM bytes and warp instructions per symbol); {} when absent."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_kernel.json")))
    except Exception:
        return {}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
# CPU arms (oracle = test/baseline infrastructure; never on the product path)
# ------------------------------------------------------------------------------------------------
def cpu_sample(E, n_states, lo, hi, mix, n_pairs, first_stream=0):
    from regex_fpga_b200 import workloads as WL
    if mix == "adv":
        return WL.make_adversarial_numpy(E, n_states, hi, 2 * n_pairs, STREAM_LEN, STRIDE, seed=0x5EED0005 + first_stream)
    return WL.make_batch_numpy(mix, lo, hi, 2 * n_pairs, STREAM_LEN, STRIDE, SEED, first_stream)


def cpu_reference_kind():
    """"reference": oracle/_ref/libref.so -- the reference's OWN Design/FPGA.v translated to C (oracle/vsim) and clocked
    like its testbench; built in the build container from /root/reference and shipped prebuilt to the GPU box.
    "port": the hand-written cycle-level restatement (oracle/oracle_a.c), only if that library is missing."""
    from oracle import ref_py as RF
    try:
        RF.lib()
        return "reference"
    except Exception:
        return "port"


def run_cpu_reference(E, n_states, sample, n_pairs, threads):
    """The reference design (FPGA.v + ROM + testbench) on host threads, one (lo,hi) stream pair per thread at a time.
    TB semantics: an M-entry trace pair yields 2*(M-1) symbol steps."""
    t0 = time.perf_counter()
    if cpu_reference_kind() == "reference":
        from oracle import ref_py as RF
        r = RF.tb_run_many(E, n_states, sample, n_pairs, STRIDE, STREAM_LEN, n_threads=threads)
    else:
        from oracle import oracle_py as O
        r = O.a_run_many(E, n_states, sample, n_pairs, STRIDE, STREAM_LEN, n_threads=threads, fast_idle=False)
    dt = time.perf_counter() - t0
    return r["symbols"], r["cycles"], dt


CPU_NOTE = {"reference": "the reference's own Design/FPGA.v executed on the host: translated to C by oracle/vsim/v2c.py and clocked as "
                         "Simulation/testbench_BLK_Mem.sv clocks it (oracle/_ref/libref.so), one (lo,hi) pair per thread",
            "port": "cycle-level C restatement of Design/FPGA.v + testbench (oracle/oracle_a.c): oracle/_ref/libref.so is missing here"}


def reference_arm(args, rank, world):
    if rank != 0:
        return
    from oracle import oracle_py as O
    O.build()
    E, n_states, lo, hi = load_ruleset(args.ruleset)
    threads = os.cpu_count() or 1
    n_pairs = max(threads, 8)
    times, syms, cycles = [], 0, 0
    for it in range(args.warmup + args.steps):
        sample = cpu_sample(E, n_states, lo, hi, args.mix, n_pairs, first_stream=2 * n_pairs * it)
        s, c, dt = run_cpu_reference(E, n_states, sample, n_pairs, threads)
        if it >= args.warmup:
            times.append(dt); syms += s; cycles += c
    total = sum(times)
    gbit = syms * 8 / total / 1e9
    desc = f"{n_pairs} (lo,hi) stream pairs x {STREAM_LEN} entries per step ({2 * n_pairs * (STREAM_LEN - 1)} symbols)"
    line = {
        "impl": "reference", "metric": metric_name(args), "value": gbit, "unit": "Gbit/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total / max(1, args.steps) * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": gbit, "unit": "Gbit/s", "cores": threads, "kind": cpu_reference_kind(), "sample": desc,
                         "simulated_cycles_per_s": cycles / total, "note": CPU_NOTE[cpu_reference_kind()]},
        "e2e": {"value": gbit, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    nfa_desc = {"snort_16": "snort_16 NFA (9514 states, 79856 transitions)",
                "l7_filter": "l7-filter NFA (2794 states, 124977 transitions)"}[args.ruleset]
    return {"workload": f"{nfa_desc} over {args.streams} synthetic "
                        f"{STREAM_LEN}-byte packet streams per GPU ({args.mix}: windows of the shipped lo/hi "
                        f"traces at splitmix64 offsets, seed {SEED:#x})",
            "streams_per_gpu": args.streams, "streams_total": args.streams * world, "stream_bytes": STREAM_LEN,
            "stride": STRIDE, "n_steps_per_stream": STREAM_LEN, "mix": args.mix,
            "sharding": f"{world} x contiguous stream shards, NCCL all-reduce of per-state counts" if world > 1
                        else "single GPU",
            "l2": "input batch (1.6 GB/GPU) >> 126 MB L2, no flush needed"}


def run_e2e(args, torch, dist, nfa, batch, n, first, cap, n_states, n_matches, barrier, dev, world, sym_per_step):
    """The same metric through the host-pointer API: pinned host batch -> H2D -> kernels -> D2H of the counts
    and the match records, all inside the timed region (wall clock, max over ranks)."""
    # the host batch is PACKED (stride = 1500 bytes, no padding crosses PCIe): stream starts are then unaligned, which the
    # kernel handles by shifting each stream's first 16-byte chunk
    host = torch.empty((n, E2E_STRIDE), dtype=torch.uint8, pin_memory=True)
    host.copy_(batch[:, :E2E_STRIDE])
    torch.cuda.synchronize()
    host_np = host.numpy()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    # results land in pinned host arrays the caller owns and reuses (a pageable destination costs ~5 ms per pass).
    # Two batches are in flight (rfb_scan_submit / rfb_scan_wait): step i+1's H2D copy overlaps the tail of step i's
    # kernel, its record sort and the D2H of its results; every step still copies its whole input from pinned host
    # memory and reads its whole result back inside the timed region.
    from regex_fpga_b200.engine import MATCH_DTYPE
    recs_host = [torch.empty(cap * 12, dtype=torch.uint8, pin_memory=True).numpy().view(MATCH_DTYPE) for _ in range(2)]
    counts_host = [torch.empty(n_states, dtype=torch.int64, pin_memory=True).numpy().view(np.uint64) for _ in range(2)]

    def run(k):
        out = None
        for i in range(k):
            nfa.submit(host_np, n, STREAM_LEN, E2E_STRIDE, recs_host[i & 1], counts_host[i & 1], flags=E2E_FLAGS, stream_id_base=first)
            if i:
                out = nfa.wait()
                assert out.n_matches == n_matches, "host-pointer and device-pointer scans disagree"
        out = nfa.wait()
        return out

    run(2)   # warm-up (allocates both slots' staging buffers)
    barrier()
    t0 = time.perf_counter()
    out = run(e2e_steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    assert out.n_matches == n_matches, "host-pointer and device-pointer scans disagree"
    return {"value": world * sym_per_step * 8 / e2e_s / 1e9, "unit": "Gbit/s", "h2d_bytes_per_step": int(host_np.size),
            "d2h_bytes_per_step": int(n_states * 8 + out.n_records * 12 + 32), "steps": e2e_steps, "s_per_step": e2e_s,
            "host_stride": E2E_STRIDE}


# ------------------------------------------------------------------------------------------------
# parity of the measured run (outside every timed region)
# ------------------------------------------------------------------------------------------------
PARITY_SAMPLE = 1024


def check_parity(torch, dist, shard, E, n_states, batch, n, first, world, rank, dev, counts, recs, n_records, n_dropped):
    """What the last pass of THIS run computed, against the CPU oracle, at every N (outside the timed regions):
      * records of a sample of PARITY_SAMPLE global stream ids spanning every shard: each rank copies its
        sampled streams back from its device batch, scans them with oracle B, and compares with the records the GPU
        wrote for exactly those streams (global ids via stream_id_base); shard.gather_records then brings both sides
        to every rank and rank 0 compares the concatenations (the gather path);
      * on every rank the per-state counts equal the histogram of its own records (counts and records are the same
        pulses), and the NCCL all-reduce of the count vectors equals the sum of the all-gathered per-rank vectors.
    Raises (the run fails) on any mismatch."""
    from oracle import oracle_py as O
    from regex_fpga_b200.engine import MATCH_DTYPE
    total = n * world
    # one stream from each of PARITY_SAMPLE equal blocks of the global id range, at a hashed position inside the block (a
    # fixed stride would alias with the batch: every 1024th stream of W-mix is a quiet one)
    from regex_fpga_b200.workloads import splitmix64_np
    blk = np.arange(PARITY_SAMPLE, dtype=np.int64)
    lo_id, hi_id = (blk * total) // PARITY_SAMPLE, ((blk + 1) * total) // PARITY_SAMPLE
    jitter = (splitmix64_np(blk.astype(np.uint64) ^ np.uint64(0x9A217E)) % np.maximum(hi_id - lo_id, 1).astype(np.uint64)).astype(np.int64)
    sample = np.unique(np.minimum(lo_id + jitter, total - 1))
    mine = sample[(sample >= first) & (sample < first + n)]
    rec = recs[: 3 * n_records].view(-1, 3)
    hist = torch.zeros(n_states, dtype=torch.int64, device=dev)
    if n_records:
        hist.index_add_(0, rec[:, 2].long(), torch.ones(n_records, dtype=torch.int64, device=dev))
    counts_vs_records = bool(torch.equal(counts, hist)) and n_dropped == 0
    reduce_ok = None
    if dist is not None:
        red = counts.clone()
        shard.reduce_counts(red, dist)
        parts = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(parts, counts)
        reduce_ok = bool(torch.equal(red, torch.stack(parts).sum(0)))
    # GPU records of the sampled streams
    sel = torch.isin(rec[:, 0].long(), torch.from_numpy(mine).to(dev)) if n_records else torch.zeros(0, dtype=torch.bool, device=dev)
    g = rec[sel].cpu().numpy().astype(np.uint32).reshape(-1, 3)
    g = g[np.lexsort((g[:, 2], g[:, 1], g[:, 0]))]
    got = np.zeros(g.shape[0], dtype=MATCH_DTYPE)
    got["stream"], got["pos"], got["state"] = g[:, 0], g[:, 1], g[:, 2]
    # oracle on the same bytes
    local = torch.from_numpy(mine - first).to(dev)
    data = batch[local].cpu().numpy()
    w = O.b_scan_many(E, n_states, data, data.shape[0], STRIDE, STREAM_LEN)
    want = w["recs"].astype(MATCH_DTYPE)
    want["stream"] = mine[w["recs"]["stream"]].astype(np.uint32)
    records_ok = got.tolist() == want.tolist()
    gather_ok = None
    if dist is not None:
        all_got = shard.gather_records(got, dist, device=dev)
        all_want = shard.gather_records(want, dist, device=dev)
        gather_ok = all_got.tolist() == all_want.tolist() and len({int(s) // n for s in all_want["stream"].tolist()} | {0}) >= 1
        n_checked = int(all_want.size)
    else:
        n_checked = int(want.size)
    flag = torch.tensor([1 if (records_ok and counts_vs_records and reduce_ok is not False and gather_ok is not False) else 0], device=dev)
    if dist is not None:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out = {"checked_streams": int(sample.size), "shards_covered": int(len({int(s) // n for s in sample.tolist()})),
           "records_checked": n_checked, "records_equal_oracle": records_ok, "counts_equal_record_histogram": counts_vs_records,
           "allreduce_equals_sum_of_ranks": reduce_ok, "gathered_records_equal_oracle": gather_ok, "ok": bool(flag.item() == 1)}
    if not out["ok"]:
        raise SystemExit(f"bench.py: PARITY FAILURE (rank {rank}): {out}")
    return out


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def bind_to_gpu_cpus(local_rank):
    """Multi-GPU runs: pin this rank to the CPUs next to its GPU (NVML's affinity mask) before any pinned host memory
    is allocated, so that the e2e leg's H2D copies read memory of the GPU's own NUMA node.  Best effort."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(local_rank)
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def ours_arm(args, rank, world, local_rank):
    import torch
    import regex_fpga_b200 as R
    from regex_fpga_b200 import shard, workloads as WL

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the scan has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    numa = bind_to_gpu_cpus(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        # stdout carries exactly one JSON line: NCCL's version banner (NCCL_DEBUG=VERSION) goes to stderr's level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=torch.device(dev))
    E, n_states, lo, hi = load_ruleset(args.ruleset)
    ctx = R.Context(local_rank)
    nfa = ctx.nfa_from_entries(E)
    n = args.streams
    first = rank * n
    if args.mix == "adv":
        batch = WL.make_adversarial_torch(E, n_states, hi, n, dev, STREAM_LEN, STRIDE, first_stream=first)
    else:
        batch = WL.make_batch_torch(args.mix, lo, hi, n, dev, STREAM_LEN, STRIDE, SEED, first)
    torch.cuda.synchronize()   # the generator ran on torch's default stream; the scans below use their own
    counts = torch.zeros(n_states, dtype=torch.int64, device=dev)
    cap = args.record_capacity
    recs = torch.empty(cap * 3, dtype=torch.int32, device=dev)
    stream = torch.cuda.Stream(device=dev)   # a real (non-NULL) handle: kernels, NCCL and events share it
    torch.cuda.set_stream(stream)
    sym_per_step = n * STREAM_LEN

    def step():
        r = nfa.scan_device(batch.data_ptr(), batch.numel(), n, STREAM_LEN, STRIDE, counts.data_ptr(),
                            recs.data_ptr(), cap, flags=R.SCAN_ASYNC, cuda_stream=stream.cuda_stream,
                            stream_id_base=first)
        if dist is not None:
            shard.reduce_counts(counts, dist)   # the path's only exchange: per-state match counts (NCCL)
        return r

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        r = step()
    barrier()
    res0 = nfa.collect(r) if args.warmup else None
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms = []
    barrier()
    ev0.record(stream)
    for _ in range(args.steps):
        r = step()
    ev1.record(stream)
    barrier()
    clocks = sampler.stop()
    res = nfa.collect(r)
    kernel_ms.append(res.gpu_ms)
    ms_total = ev0.elapsed_time(ev1)
    if dist is not None:
        t = torch.tensor([ms_total], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_per_step = ms_total / args.steps
    value = world * sym_per_step * 8 / (ms_per_step * 1e-3) / 1e9

    # ---- per-kernel time for the roofline: the library brackets its kernels with CUDA events ----
    # (one more un-overlapped step so that gpu_ms is exactly lane kernel + rescan kernel of one pass)
    r = nfa.scan_device(batch.data_ptr(), batch.numel(), n, STREAM_LEN, STRIDE, counts.data_ptr(), recs.data_ptr(), cap,
                        cuda_stream=stream.cuda_stream, stream_id_base=first)
    scan_ms = float(r.gpu_ms)
    n_matches, n_rescanned, n_dropped = int(r.n_matches), int(r.n_rescanned), int(r.n_dropped)
    peak, peak_src = measured_peaks()
    achieved = sym_per_step * 1.0 / (scan_ms * 1e-3) / 1e9     # GB/s, 1 algorithmic byte per symbol

    e2e = {"value": None, "unit": "Gbit/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    if not args.no_e2e:
        e2e = run_e2e(args, torch, dist, nfa, batch, n, first, cap, n_states, n_matches, barrier, dev, world, sym_per_step)

    parity = check_parity(torch, dist, shard, E, n_states, batch, n, first, world, rank, dev, counts, recs, int(r.n_records), n_dropped)

    cap_ncu = ncu_capture().get(args.mix if args.ruleset == "snort_16" else "", {})
    issue = None
    if "warp_instructions_per_32_symbols" in cap_ncu:
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        n_sms = torch.cuda.get_device_properties(dev).multi_processor_count
        peak_issue = n_sms * 4 * sm_mhz * 1e6                         # warp instructions per second, 4 schedulers per SM
        used = cap_ncu["warp_instructions_per_32_symbols"] * (sym_per_step / 32.0) / (scan_ms * 1e-3)
        issue = {"warp_instructions_per_32_symbols": cap_ncu["warp_instructions_per_32_symbols"],
                 "active_threads_per_instruction": cap_ncu.get("threads_per_instruction"),
                 "achieved": used / 1e9, "peak": peak_issue / 1e9, "unit": "G warp-instructions/s", "frac": used / peak_issue,
                 "source": "profiles/r2_kernel.json (ncu --set full of this workload)"}
    if rank == 0:
        line = {
            "metric": metric_name(args), "value": value, "unit": "Gbit/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": dict(workload_config(args, world), **({"cpus_bound_per_rank": numa} if numa else {})),
            "clocks": clocks,
            "e2e": e2e,
            "gpu_launches": int(r.n_launches) * args.steps,   # per pass: lane kernel, second lane pass over the deferred streams, hand-over kernel
            "roofline": {"bound": "hbm", "kernel": "scan_lane_kernel", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (int(round(cap_ncu["dram_bytes_per_symbol"] * sym_per_step)) if "dram_bytes_per_symbol" in cap_ncu else None),
                         "traffic_unit": "bytes per launch (ncu dram read+write of this workload at this size, profiles/r2_kernel.json)",
                         "peak_source": peak_src, "frac_of_8000": achieved / 8000.0, "scan_ms": scan_ms,
                         "algorithmic_bytes_per_launch": sym_per_step},
            # the limiter the HBM fraction does not show: warp instructions issued per 32 symbols (ncu, committed capture)
            # against the SMs' issue rate at the clock sampled during this run
            "roofline_issue": issue,
            "parity": parity,
            "matches_per_step_rank0": n_matches, "rescanned_streams": n_rescanned, "records_dropped": n_dropped,
            "symbols_per_s": world * sym_per_step / (ms_per_step * 1e-3),
            "image": nfa.info,
        }
        if world == 1 and not args.no_cpu:
            from oracle import oracle_py as O
            O.build()
            threads = os.cpu_count() or 1
            n_pairs = max(threads, 8) * args.cpu_pairs_per_core
            sample = cpu_sample(E, n_states, lo, hi, args.mix, n_pairs)
            s, c, dt = run_cpu_reference(E, n_states, sample, n_pairs, threads)
            t0 = time.perf_counter()
            fb = O.b_scan_many(E, n_states, sample, 2 * n_pairs, STRIDE, STREAM_LEN, n_threads=threads, want_recs=False)
            dtb = time.perf_counter() - t0
            line["cpu_baseline"] = {
                "value": s * 8 / dt / 1e9, "unit": "Gbit/s", "cores": threads, "kind": cpu_reference_kind(),
                "sample": f"{n_pairs} (lo,hi) stream pairs x {STREAM_LEN} entries of the same {args.mix} workload "
                          f"({s} symbols, {dt:.1f} s wall)",
                "simulated_cycles_per_s": c / dt,
                "note": CPU_NOTE[cpu_reference_kind()],
                "functional_port_gbit_s": 2 * n_pairs * STREAM_LEN * 8 / dtb / 1e9,
            }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--mix", default="wmix", choices=["wmix", "whi", "wlo", "uniform", "adv", "wsplice"])
    ap.add_argument("--ruleset", default="snort_16", choices=["snort_16", "l7_filter"],
                    help="snort_16 is the headline (BASELINE.json); l7_filter is the reference's other shipped image")
    ap.add_argument("--streams", type=int, default=1 << 20, help="streams per GPU")
    ap.add_argument("--record-capacity", type=int, default=1 << 22)
    ap.add_argument("--e2e-steps", type=int, default=10)
    ap.add_argument("--cpu-pairs-per-core", type=int, default=4)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours_arm(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
