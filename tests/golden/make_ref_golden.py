#!/usr/bin/env python
"""Regenerates tests/golden/ref_vsim.json by EXECUTING THE REFERENCE'S OWN VERILOG.

Run in the build container (needs /root/reference; takes about two minutes):
    python tests/golden/make_ref_golden.py
What runs: Design/FPGA.v, unmodified and read where it lies, translated to C by oracle/vsim/v2c.py
(`make -C oracle _ref`) and clocked by oracle/vsim/tb_driver.c the way Simulation/testbench_BLK_Mem.sv clocks it
(M = 200000 trace entries, TB:71; size_range = the ruleset's state count, TB:20), on the reference's own
.coe / .mem files as decoded into tests/golden/<ruleset>.npz.  What is stored, per ruleset: the testbench's
printout -- non-zero match_count / match_count_2 entries (10-bit counters) and "Total no. cycles" -- plus the
match pulses as (pos, state) events with SHA-256 digests in the canonical text form of SURVEY.md Appendix C,
the cycle total of the first 2000 trace entries and a digest of the per-edge port trace of that prefix
(i, input_char_flag, accepting_match_flag(_2), state, rd_address after every posedge).
Both x-fills of the 2-state model (see v2c.py) must agree on everything the testbench observes.
"""
import hashlib
import json
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import ref_py as RF  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
TB_M = 200000       # testbench_BLK_Mem.sv:71
PREFIX_M = 2000
PORT_MASK = np.uint64((1 << 32) - 1)    # i | flags | state: everything but rd_address (not reset, x until first written)


def digest_counts(counts):
    s = "".join(f"{i} {int(counts[i])}\n" for i in np.nonzero(counts)[0])
    return hashlib.sha256(s.encode()).hexdigest()[:32]


def digest_events(recs):
    s = "".join(f"{int(p)} {int(st)}\n" for p, st in zip(recs["pos"], recs["state"]))
    return hashlib.sha256(s.encode()).hexdigest()[:32]


def summarise(run, stream, counts, mc):
    ev = run["recs"][run["recs"]["stream"] == stream]
    nz = np.nonzero(counts)[0]
    return {"n_matches": int(len(ev)), "counts": {str(int(i)): int(counts[i]) for i in nz},
            "match_count_10bit": {str(int(i)): int(mc[i]) for i in np.nonzero(mc)[0]},
            "events": [[int(p), int(s)] for p, s in zip(ev["pos"], ev["state"])],
            "counts_digest": digest_counts(counts), "events_digest": digest_events(ev)}


def run_ruleset(name, out):
    z = np.load(os.path.join(OUT, name + ".npz"))
    E, size, lo, hi = z["entries"], int(z["n_states"]), z["lo"], z["hi"]
    full = RF.tb_run(E, size, lo, hi, TB_M, xfill=0)
    pre = {x: RF.tb_run(E, size, lo, hi, PREFIX_M, xfill=x, trace_cap=1 << 25) for x in (0, 1)}
    assert pre[0]["cycles"] == pre[1]["cycles"]
    assert np.array_equal(pre[0]["trace"] & PORT_MASK, pre[1]["trace"] & PORT_MASK), "x-fill changes what the testbench sees"
    assert np.array_equal(pre[0]["trace"][1:], pre[1]["trace"][1:]), "rd_address depends on the x-fill after the reset edge"
    assert pre[0]["recs"].tobytes() == pre[1]["recs"].tobytes()
    out[name] = {
        "n_states": size, "trace_entries": TB_M,
        "total_cycles": int(full["cycles"]),
        "lo": summarise(full, 0, full["counts1"], full["mc1"]),
        "hi": summarise(full, 1, full["counts2"], full["mc2"]),
        "prefix": {"trace_entries": PREFIX_M, "cycles": int(pre[0]["cycles"]), "n_matches": int(pre[0]["n_recs"]),
                   "port_trace_sha256": hashlib.sha256((pre[0]["trace"] & PORT_MASK).tobytes()).hexdigest(),
                   "full_trace_sha256_xfill0": hashlib.sha256(pre[0]["trace"].tobytes()).hexdigest()},
    }
    print(name, size, out[name]["total_cycles"], out[name]["lo"]["n_matches"], out[name]["hi"]["n_matches"], flush=True)


def main():
    RF.build()
    res = {}
    th = [threading.Thread(target=run_ruleset, args=(n, res)) for n in ("snort_16", "l7_filter")]
    [t.start() for t in th]
    [t.join() for t in th]
    src = os.path.join(RF.REFERENCE, "Design", "FPGA.v")
    doc = {"generated_by": "tests/golden/make_ref_golden.py (oracle/vsim/v2c.py + tb_driver.c on the reference's Design/FPGA.v)",
           "module": RF.lib().ref_tb_module().decode(),
           "fpga_v_sha256": hashlib.sha256(open(src, "rb").read()).hexdigest(),
           "rulesets": res}
    with open(os.path.join(OUT, "ref_vsim.json"), "w") as f:
        json.dump(doc, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
