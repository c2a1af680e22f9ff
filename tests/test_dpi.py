"""The DPI-C shim of INTEGRATION.md (tools/dpi/rfb_dpi.c): compiles with plain gcc against the public header, links to
librfb200.so, and -- called the way the imported SystemVerilog functions would call it -- prints the testbench's report."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "regex_fpga_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "dpi_caller")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tools", "dpi", "rfb_dpi.c"), os.path.join(ROOT, "tests", "dpi_caller.c"),
                           "-o", exe, "-L", LIBDIR, "-lrfb200", "-Wl,-rpath," + LIBDIR])
    return exe


def inputs(tmp_path, rs, M):
    import regex_fpga_b200 as R
    coe, lo, hi = tmp_path / "n.coe", tmp_path / "lo.bin", tmp_path / "hi.bin"
    R.coe_write(coe, rs.entries, 1)
    np.ascontiguousarray(rs.lo[:M]).tofile(lo)
    np.ascontiguousarray(rs.hi[:M]).tofile(hi)
    return [str(coe), str(rs.n_states), str(lo), str(hi), str(M)]


def test_dpi_shim_compiles_and_links(tmp_path, l7):
    exe = build(tmp_path)
    assert subprocess.check_output([exe] + inputs(tmp_path, l7, 100) + ["--host-only"], text=True).strip() == "host only"


@pytest.mark.gpu
def test_dpi_shim_reproduces_the_testbench_report(tmp_path, l7):
    from oracle import oracle_py as O
    exe = build(tmp_path)
    M = 30000
    out = subprocess.check_output([exe] + inputs(tmp_path, l7, M), text=True).strip().split("\n")
    out = [ln for ln in out if not ln.startswith("NCCL version")]
    a = O.a_run(l7.entries, l7.n_states, l7.lo, l7.hi, M, fast_idle=True)
    want = []
    for label, mc in (("match_count", a["mc1"]), ("match_count_2", a["mc2"])):
        want += [f"{label}[{p}] = {int(mc[p])}" for p in np.nonzero(mc)[0][::-1]]
    want.append(f"Total no. cycles: {a['cycles']}")
    assert out == want
