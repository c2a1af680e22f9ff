#!/usr/bin/env python
"""ncu_summary.py REPORT.ncu-rep [n_symbols]: key raw metrics of the first kernel in an ncu report and the per-source-line
instruction table (needs -lineinfo and --import-source on).  Dev tool; its output is what gets committed under profiles/."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
nsym = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, vals = rows[0], rows[-1]
M = dict(zip(hdr, vals))
keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sector_hit_rate.pct",
        "smsp__average_warp_latency_per_inst_issued.ratio", "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic"]
keys += [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
print("# kernel:", M.get("Kernel Name"))
for k in keys:
    if k in M:
        print(f"{k}\t{M[k]}")
if nsym:
    print(f"warp_instructions_per_32_symbols\t{float(M['smsp__inst_executed.sum']) / (nsym / 32):.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
def num(x):
    try: return float(x)
    except ValueError: return 0.0
lines, fname = [], "?"
h = None
for r in rows:
    if len(r) >= 2 and r[0] == "File Name": fname = r[1].split("/")[-1]; continue
    if len(r) > 8 and r[0] == "Line No": h = r; ie = h.index("Instructions Executed"); te = h.index("Thread Instructions Executed"); ss = h.index("Warp Stall Sampling (All Samples)"); continue
    if h is None or len(r) <= te: continue
    n = num(r[ie])
    if n <= 0: continue
    lines.append((n, num(r[te]), num(r[ss]), fname, r[0], r[1].strip()))
tot = sum(l[0] for l in lines) or 1
stot = sum(l[2] for l in lines) or 1
print(f"# source lines by warp instructions executed (total {tot:.0f})")
print("# file:line  instr%  threads/instr  stall-samples%  source")
for n, t, s_, f, ln, text in sorted(lines, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 45]:
    print(f"{f}:{ln:>4} {100 * n / tot:6.2f} {t / n:6.1f} {100 * s_ / stot:6.2f}  {text[:120]}")
