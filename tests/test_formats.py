"""The reference's two text formats: product parser/writer (librfb200.so, host-only entry points)
against the oracle's independent decoder and against the golden decode of the shipped files."""
import os

import numpy as np
import pytest

import regex_fpga_b200 as R
from oracle import oracle_py as O

REF = "/root/reference"


@pytest.mark.parametrize("style", [0, 1])
def test_coe_roundtrip_both_styles(tmp_path, snort, l7, style):
    rs = snort if style == 0 else l7
    p = tmp_path / "img.coe"
    R.coe_write(p, rs.entries, style)
    txt = p.read_text()
    assert txt.startswith("memory_initialization_radix=16;\nmemory_initialization_vector=")
    if style == 0:   # CSR_BlockMem_snort_16.coe: newline separated, no terminator, no final newline
        assert not txt.rstrip("\n").endswith(";") and txt.count("\n") == rs.entries.size // 4
    else:            # CSR_BlockMem.coe: one line, blank separated, ';' terminated
        assert txt.rstrip("\n").endswith(";")
    assert np.array_equal(R.coe_parse(p), rs.entries)
    assert np.array_equal(O.coe_parse(str(p)), rs.entries)
    assert R.coe_detect_size(rs.entries) == rs.n_states == O.detect_size(rs.entries)


def test_coe_first_line_layout(tmp_path, snort):
    """Leftmost 8 hex digits = entry 4*line+0 (rd_bus[127:96] = cache[0], Design/FPGA.v:884)."""
    p = tmp_path / "a.coe"
    p.write_text("memory_initialization_radix=16;\nmemory_initialization_vector=\n"
                 "00000000000001280000024a00000256,\n000002580000025a0000025c00000260;")
    e = R.coe_parse(p)
    assert e.tolist() == [0, 0x128, 0x24A, 0x256, 0x258, 0x25A, 0x25C, 0x260]
    assert snort.entries[:8].tolist() == e.tolist()      # CSR_BlockMem_snort_16.coe:2-3


def test_coe_errors(tmp_path):
    bad = tmp_path / "bad.coe"
    bad.write_text("memory_initialization_radix=16;\nmemory_initialization_vector=0000000000000128\n")
    with pytest.raises(R.RfbError) as e:
        R.coe_parse(bad)
    assert e.value.code == -3
    bad.write_text("memory_initialization_radix=10;\nmemory_initialization_vector=1 2 3;")
    with pytest.raises(R.RfbError):
        R.coe_parse(bad)
    with pytest.raises(R.RfbError) as e:
        R.coe_parse(tmp_path / "missing.coe")
    assert e.value.code == -2


def test_mem_roundtrip(tmp_path, snort):
    p = tmp_path / "t.mem"
    data = np.concatenate([np.arange(256, dtype=np.uint8), snort.hi[:1000]])
    R.trace_write_mem(p, data)
    lines = p.read_text().split("\n")
    assert lines[0] == "0" and lines[15] == "f" and lines[16] == "10" and lines[255] == "ff"   # no leading zero
    assert np.array_equal(R.trace_load_mem(p), data)
    assert np.array_equal(O.mem_parse(str(p)), data)
    p.write_text("c6\nc6\n7f\n1ff\n")
    with pytest.raises(R.RfbError):
        R.trace_load_mem(p)


def test_tb_steps():
    assert R.tb_steps(200000) == 199999 and R.tb_steps(1) == 0 and R.tb_steps(0) == 0


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout not present (GPU box)")
@pytest.mark.parametrize("name,coe,lo,hi", [
    ("snort_16", "Block_Mem/CSR_BlockMem_snort_16.coe", "Simulation/input_trace_lo_snort_16.mem",
     "Simulation/input_trace_hi_snort_16.mem"),
    ("l7_filter", "Block_Mem/CSR_BlockMem.coe", "Simulation/input_trace_lo_l-7_filter.mem",
     "Simulation/input_trace_hi_l-7_filter.mem")])
def test_shipped_files_decode_to_golden(name, coe, lo, hi, snort, l7):
    rs = snort if name == "snort_16" else l7
    assert np.array_equal(R.coe_parse(os.path.join(REF, coe)), rs.entries)
    assert np.array_equal(R.trace_load_mem(os.path.join(REF, lo)), rs.lo)
    assert np.array_equal(R.trace_load_mem(os.path.join(REF, hi)), rs.hi)
    assert R.coe_detect_size(rs.entries) == rs.n_states


def test_snort_boundary_line(snort):
    """CSR_BlockMem_snort_16.coe:2380 = row_ptr[9512], row_ptr[9513], row_ptr[9514]=nnz, transition 0."""
    e = snort.entries
    assert e[9512:9516].tolist() == [0x137EE, 0x137F0, 0x137F0, 0x00000001]
    assert e[9514] == 79856 and e.size == 22343 * 4
