"""BASELINE configs[0]: n = 2 encode + decode round trip of a 1 MiB synthetic skewed-byte stream on the CPU, with the tables
coming from the UNMODIFIED reference (histogram() n_ary_huffman.c:461, huffman() :1161, convert_lengths_to_encode_table()
:1382 through oracle/_ref, built with -DNDEBUG) and the payload from the oracle's packer and decoder -- the reference has
neither.  The restatement must agree with the reference on every table."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def test_config1_one_mib_binary_roundtrip_through_the_reference_build():
    from oracle import pyoracle as O
    O.build()
    if not O.have_ref():
        pytest.skip("oracle/_ref is not built here (needs /root/reference once)")
    from data_compression_b200 import synth
    thr, base = synth.zipf_7bit_spec()              # ranks 1..126: the unmodified histogram() prints on bytes > 126
    data = synth.host_stream(1 << 20, synth.SEED_BASE + 0, thr, base)
    assert data.size == 1 << 20 and data.min() >= 1 and data.max() <= 126
    hist = O.ref_histogram(data.tobytes())                                   # unmodified histogram()
    assert hist[258] == 0 and int(hist.sum()) == data.size
    assert np.array_equal(hist.astype(np.uint64), O.histogram_u8(data)[:259])
    lengths = O.ref_huffman(hist, 2)                                         # unmodified huffman(n = 2)
    elen, evalue = O.ref_convert_lengths_to_encode_table(lengths, 2)         # unmodified canonical values
    o_len, o_elen, o_eval, st = O.build_tables(hist.astype(np.uint64), 2)    # the restatement agrees
    assert st == 0 and np.array_equal(o_len, lengths) and np.array_equal(o_elen, elen) and np.array_equal(o_eval, evalue)
    payload, bits = O.pack(data, elen, evalue, 1)
    assert bits == int((hist.astype(np.int64) * lengths).sum())
    assert payload.size == (bits + 7) // 8 < data.size                       # it does compress
    back = O.unpack(payload, 0, bits, lengths, 2, data.size)
    assert np.array_equal(back, data)
    # a flipped bit must not go unnoticed: either an error or different symbols
    bad = payload.copy()
    bad[bad.size // 2] ^= 0x10
    try:
        wrong = O.unpack(bad, 0, bits, lengths, 2, data.size)
        assert not np.array_equal(wrong, data)
    except Exception:
        pass
