"""The drop-in boundary is a C ABI: a plain-C host program (tests/c_abi_example.c, the one shown in INTEGRATION.md)
compiled with gcc against include/regex_fpga_b200.h and linked to librfb200.so -- no Python in that process."""
import os
import subprocess

import numpy as np
import pytest

import regex_fpga_b200 as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "regex_fpga_b200", "lib")


def build(tmp_path):
    exe = str(tmp_path / "c_abi_example")
    subprocess.check_call(["gcc", "-std=c11", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "c_abi_example.c"), "-o", exe, "-L", LIBDIR, "-lrfb200",
                           "-Wl,-rpath," + LIBDIR])
    return exe


def write_inputs(tmp_path, rs, entries):
    coe, lo, hi = tmp_path / "n.coe", tmp_path / "lo.mem", tmp_path / "hi.mem"
    R.coe_write(coe, rs.entries, 0)
    R.trace_write_mem(lo, rs.lo[:entries])
    R.trace_write_mem(hi, rs.hi[:entries])
    return [str(coe), str(lo), str(hi), str(entries)]


def test_header_is_valid_c_and_links(tmp_path, l7):
    """CPU box: the header compiles as C11 with -Werror, every used symbol resolves, the host-only entry points run."""
    exe = build(tmp_path)
    out = subprocess.check_output([exe] + write_inputs(tmp_path, l7, 100) + ["--host-only"], text=True)
    assert out.strip() == "size_range = 2794"                         # testbench_BLK_Mem.sv:20


@pytest.mark.gpu
def test_c_host_program_reproduces_testbench_report(tmp_path, snort, expected):
    exe = build(tmp_path)
    M = 20000
    out = subprocess.check_output([exe] + write_inputs(tmp_path, snort, M), text=True).strip().split("\n")
    out = [ln for ln in out if not ln.startswith("NCCL version")]     # NCCL_DEBUG=VERSION makes the library announce itself on stdout
    assert out[0] == "size_range = 9514"
    from oracle import oracle_py as O
    a = O.a_run(snort.entries, snort.n_states, snort.lo, snort.hi, M, fast_idle=True)
    want = []
    for label, mc in (("match_count", a["mc1"]), ("match_count_2", a["mc2"])):
        want += [f"{label}[{p}] = {int(mc[p])}" for p in np.nonzero(mc)[0][::-1]]
    want.append(f"Total no. cycles: {a['cycles']}")
    assert out[1:] == want
