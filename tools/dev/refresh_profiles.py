#!/usr/bin/env python
"""refresh_profiles.py [SRC_DIR]: rebuild the round-2 evidence under profiles/ from what
tools/dev/round_end_measurements.sh left in SRC_DIR (default gpurun_out/final): bench lines, the launch list, config 5,
and per mix the raw ncu metrics (.tsv), the per-source-line table and the figures bench.py reads (r2_kernel.json).
Runs here (ncu -i reads reports without a GPU)."""
import csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "..")
SRC = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "final")
DST = os.path.join(ROOT, "profiles")
STREAMS, STREAM_BYTES = 1 << 20, 1500

for f in sorted(os.listdir(SRC)):
    if f.startswith("r2_bench_n1") or f == "r2_launches_bench_steps2_warmup3.csv":
        shutil.copy(os.path.join(SRC, f), os.path.join(DST, f))
if os.path.exists(os.path.join(SRC, "config5.txt")):
    shutil.copy(os.path.join(SRC, "config5.txt"), os.path.join(DST, "r2_config5.txt"))

kernel = {}
for mix in ("wmix", "whi", "wlo"):
    rep = os.path.join(SRC, f"prof_r2_{mix}.ncu-rep")
    if not os.path.exists(rep):
        continue
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[-1]
    tsv = os.path.join(DST, f"r2_scan_lane_ncu_metrics_{mix}.tsv")
    with open(tsv, "w") as fh:
        fh.write(f"# ncu --set full --clock-control none, scan_lane_kernel<1,16>, python bench.py --steps 1 --warmup 3 --no-cpu "
                 f"--no-e2e --mix {mix} (1 Mi streams x 1500 B), 4th launch\n")
        for h, u, v in zip(hdr, units, vals):
            fh.write(f"{h}\t{u}\t{v}\n")
    M = dict(zip(hdr, vals))
    U = dict(zip(hdr, units))

    def val(k, to=None):
        x = float(M[k].replace(",", ""))
        u = U.get(k, "").split("/")[0]
        if to == "byte":
            x *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]
        if to == "ms":
            x *= {"ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}[u]
        return x
    nsym = STREAMS * STREAM_BYTES
    rd, wr = val("dram__bytes_read.sum", "byte"), val("dram__bytes_write.sum", "byte")
    stalls = {}
    for k in hdr:
        if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio"):
            v = round(val(k), 2)
            if v >= 0.3:
                stalls[k[len("smsp__average_warps_issue_stalled_"):-len("_per_issue_active.ratio")]] = v
    kernel[mix] = {
        "dram_bytes_per_symbol": (rd + wr) / nsym, "dram_read_bytes": rd, "dram_write_bytes": wr,
        "warp_instructions_per_32_symbols": val("smsp__inst_executed.sum") / (nsym / 32),
        "threads_per_instruction": val("smsp__thread_inst_executed_per_inst_executed.ratio"),
        "issue_active_pct": val("smsp__issue_active.avg.pct_of_peak_sustained_active"),
        "lsu_wavefronts_pct_of_peak": val("l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
        "shared_bank_conflict_wavefronts": val("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"),
        "l1_hit_pct": val("l1tex__t_sector_hit_rate.pct"),
        "kernel_ms_under_ncu": val("gpu__time_duration.sum", "ms"),
        "registers_per_thread": int(val("launch__registers_per_thread")),
        "dynamic_shared_bytes": val("launch__shared_mem_per_block_dynamic", "byte"),
        "stalls_per_issue": stalls,
        "capture": f"profiles/r2_scan_lane_ncu_metrics_{mix}.tsv", "streams": STREAMS, "stream_bytes": STREAM_BYTES,
    }
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "dev", "ncu_summary.py"), rep, str(nsym)],
                         capture_output=True, text=True).stdout
    open(os.path.join(DST, f"r2_scan_lane_hot_lines_{mix}.txt"), "w").write(out)
    print(mix, {k: kernel[mix][k] for k in ("kernel_ms_under_ncu", "warp_instructions_per_32_symbols", "threads_per_instruction",
                                            "issue_active_pct", "dram_bytes_per_symbol", "l1_hit_pct")}, stalls)
if kernel:
    json.dump(kernel, open(os.path.join(DST, "r2_kernel.json"), "w"), indent=1)
