"""CPU: the C-ABI library loads and exports every symbol include/dc_b200.h declares; the product package
never reaches into oracle/ and fails loudly without a device."""
import ctypes
import os
import re
import subprocess

import pytest

from conftest import ROOT


def _header_symbols():
    text = open(os.path.join(ROOT, "include", "dc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(dc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()
    from data_compression_b200 import _lib
    return _lib


def test_library_exports_every_declared_symbol(built):
    syms = _header_symbols()
    assert len(syms) >= 25
    L = ctypes.CDLL(built.LIB_PATH)
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing
    declared = {name for name, _, _ in built.SYMBOLS}
    assert declared == set(syms), declared ^ set(syms)


def test_table_struct_layout(built):
    # offsets the kernels and the Python mirror must agree on
    T = built.HuffTableStruct
    assert ctypes.sizeof(T) == 14304 + 2 * 6564 * 4 + 257 * 16 * 2 + 8 + 2 * (1 << 14)
    assert T.total_symbols.offset == 40 and T.lengths.offset == 56 and T.enc64.offset % 8 == 0
    assert built.lib().dc_version().startswith(b"dc_b200")
    assert built.lib().dc_status_string(-5) == b"corrupt bitstream"


def test_no_cpu_fallback(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-device behaviour")
    assert built.lib().dc_device_count() < 0 or built.lib().dc_device_count() == 0
    import numpy as np
    from data_compression_b200 import DcError, hostapi
    with pytest.raises(DcError) as e:
        hostapi.histogram_u8(np.zeros(16, dtype=np.uint8))
    assert e.value.status == built.DC_ERR_CUDA
    import data_compression_b200 as dc
    with pytest.raises(RuntimeError):
        dc.histogram(torch.zeros(16, dtype=torch.uint8))


def test_product_never_touches_oracle():
    pkg = os.path.join(ROOT, "data_compression_b200")
    hits = subprocess.run(["grep", "-rIl", "-E", r"oracle|pyoracle|liboracle|_ref/", pkg, os.path.join(ROOT, "include")],
                          capture_output=True, text=True).stdout.split()
    hits = [h for h in hits if not h.endswith((".so", ".o"))]
    assert not hits, hits
    # the shared library links nothing from oracle/
    from data_compression_b200 import _lib
    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_synthetic_stream_is_deterministic():
    import numpy as np
    from data_compression_b200 import synth
    thr, base = synth.zipf_bytes_spec()
    a = synth.host_stream(100000, synth.SEED_BASE + 3, thr, base)
    assert a.min() >= 1 and a.max() <= 255
    b = synth.host_stream(50000, synth.SEED_BASE + 3, thr, base, start=50000)
    assert np.array_equal(a[50000:], b)
    p = np.bincount(a, minlength=256)[1:4] / a.size
    assert abs(p[0] - 0.2066) < 0.01  # Zipf(1.1) over 255 ranks
    thr4, base4 = synth.zipf_nybble_spec()
    s = synth.host_stream(4096, 1, thr4, base4)
    assert s.max() <= 15
