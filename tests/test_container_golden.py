"""The netstring container against the reference's OWN writer (static compress(), n_ary_huffman.c:1688-1815), CPU only:
golden buffers in tests/golden/container.json were produced by running the unmodified function (oracle/ref_harness_huff.c,
tests/golden/make_golden.py).  Under -DNDEBUG the reference always ends in the raw pass-through block; the host-side writer
must produce exactly those bytes when it falls back to the raw block, and the reader must take them."""
import ctypes
import json
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFAPI = os.path.join(ROOT, "data_compression_b200", "libdc_b200_refapi.so")


@pytest.fixture(scope="module")
def L():
    if not os.path.exists(REFAPI):
        import sys
        sys.path.insert(0, ROOT)
        import __graft_entry__ as g
        g.build()
    lib = ctypes.CDLL(REFAPI)
    lib.dc_container_compress.restype = ctypes.c_size_t
    lib.dc_container_compress.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t,
                                          ctypes.c_char_p, ctypes.c_size_t]
    lib.dc_container_decompress.restype = ctypes.c_size_t
    lib.dc_container_decompress.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    return lib


def _cases():
    return json.load(open(os.path.join(ROOT, "tests", "golden", "container.json")))["cases"]


def test_raw_block_is_the_reference_writers(L):
    lengths = (ctypes.c_int * 259)()
    for case in _cases():
        text = bytes.fromhex(case["text"])
        buf = bytes.fromhex(case["buffer"])
        ref_block = buf[: buf.index(b"\x00")]                 # sprintf's NUL ends what the reference wrote last
        assert ref_block == b"%d:\n\n" % (len(text) + 2) + text + b","   # (:1811) no line feed behind the comma
        out = ctypes.create_string_buffer(len(text) + 128)
        # radix 10 has no payload form, so the writer takes the raw block without touching a GPU
        n = L.dc_container_compress(10, lengths, text, len(text), out, len(text) + 127)
        assert out.raw[:n] == ref_block
        back = ctypes.create_string_buffer(len(text) + 2)
        assert L.dc_container_decompress(10, ref_block, len(ref_block), back, len(text) + 2) == len(text)
        assert back.raw[: len(text)] == text
        # ... and the form with the line feed the reference's reader asserts (:2049)
        assert L.dc_container_decompress(10, ref_block + b"\n", len(ref_block) + 1, back, len(text) + 2) == len(text)


@pytest.mark.parametrize("blob", [
    b"", b"5", b"5:", b":\n\nabc,", b"-5:\n\nabc,", b" 5:\n\nabc,", b"+5:\n\nabc,", b"0x5:\n\nabc,", b"5:\n\nabc", b"5:\n\nabcd",
    b"18446744073709551615:\n\nabc,", b"99999999999999999999999:\n\nabc,", b"4:\n\nabc,", b"5:\nQabc,", b"5:x\nabc,",
    b"3:\nX1,", b"265:\nX258:" + b"1" * 258 + b"G,", b"7:\nZ1 8\nA,", b"1:\n,",
])
def test_reader_rejects_malformed_input(L, blob):
    """lengths that wrap, signs, white space, unterminated input, headers that run off the buffer."""
    back = ctypes.create_string_buffer(64)
    got = L.dc_container_decompress(2, blob, len(blob), back, 64)
    if blob == b"":
        assert got == 0
    else:
        assert got == ctypes.c_size_t(-1).value, blob


def test_writer_checks_its_arguments(L):
    lengths = (ctypes.c_int * 259)()
    out = ctypes.create_string_buffer(64)
    assert L.dc_container_compress(10, lengths, b"x" * 100, 100, out, 63) == ctypes.c_size_t(-1).value   # no room for the raw block
    lengths[65] = -3   # a negative length must not index the digit string: falls back to the raw block
    big = ctypes.create_string_buffer(8192)
    n = L.dc_container_compress(2, lengths, b"A" * 10, 10, big, 8191)
    assert big.raw[:n] == b"12:\n\n" + b"A" * 10 + b","
