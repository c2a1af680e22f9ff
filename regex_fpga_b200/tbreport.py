"""Testbench-compatible report from GPU results.

    python -m regex_fpga_b200.tbreport --coe CSR_BlockMem.coe --lo input_trace_lo.mem --hi input_trace_hi.mem
                                        [--entries 200000] [--size N] [--device 0] [--no-cycles]

Prints what Simulation/testbench_BLK_Mem.sv prints (TB:75-84): the non-zero `match_count[p]` of stream 1 (lo
trace, input_char) and `match_count_2[p]` of stream 2 (hi trace, input_char_2) in descending p -- `foreach` over
`[size_range-1:0]` -- with the testbench's 10-bit counters, then `Total no. cycles`.  Field widths of `%d` are
simulator dependent; compare parsed pairs, not text.  Everything is computed on the GPU.
"""
import argparse

import numpy as np

from . import engine as R


def tb_report(nfa, lo, hi, trace_entries=200000, with_cycles=True):
    """Returns (lines, per-stream counters) for an M-entry testbench run."""
    lo = np.ascontiguousarray(lo[:trace_entries], dtype=np.uint8)
    hi = np.ascontiguousarray(hi[:trace_entries], dtype=np.uint8)
    if lo.size < trace_entries or hi.size < trace_entries:
        raise ValueError("traces are shorter than --entries")
    res = nfa.scan(np.stack([lo, hi]), 2, n_steps=R.tb_steps(trace_entries), stride=trace_entries,
                   record_capacity=1 << 22, flags=R.SCAN_SORT_RECORDS)
    if res.n_dropped:
        raise RuntimeError("record buffer too small")
    lines, counters = [], []
    for stream, name in ((0, "match_count"), (1, "match_count_2")):
        mc = np.bincount(res.records["state"][res.records["stream"] == stream], minlength=nfa.n_states)
        mc = mc & 0x3FF                                     # logic [9:0] match_count (TB:21-22)
        counters.append(mc)
        for p in np.nonzero(mc)[0][::-1]:                   # descending index (TB:75-81)
            lines.append(f"{name}[{int(p)}] = {int(mc[p])}")
    if with_cycles:
        lines.append(f"Total no. cycles: {nfa.fpga_cycles(lo, hi, trace_entries)}")
    return lines, counters


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--coe", required=True)
    ap.add_argument("--lo", required=True)
    ap.add_argument("--hi", required=True)
    ap.add_argument("--entries", type=int, default=200000, help="trace entries the testbench consumes (TB:71)")
    ap.add_argument("--size", type=int, default=-1, help="size_range (TB:20); default: derived from the image")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--no-cycles", action="store_true")
    a = ap.parse_args(argv)
    with R.Context(a.device) as ctx:
        nfa = ctx.load_coe(a.coe, a.size)
        lines, _ = tb_report(nfa, R.trace_load_mem(a.lo), R.trace_load_mem(a.hi), a.entries, not a.no_cycles)
    print("\n".join(lines))


if __name__ == "__main__":
    main()
