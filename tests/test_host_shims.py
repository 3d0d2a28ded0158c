"""The C host side (data_compression_b200/host): refapi library exports the reference's function names, and
the CLI shims named after the reference binaries behave like them -- `Successful test.` per round trip on a
GPU box, a loud abort (never a CPU fallback) without a device."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HOST = os.path.join(ROOT, "data_compression_b200", "host")
REFAPI = os.path.join(ROOT, "data_compression_b200", "libdc_b200_refapi.so")
REF_NAMES = ["histogram", "huffman", "convert_lengths_to_encode_table", "represent_items_with_codes",
             "decode_items_with_codes", "write_nybble", "nybble_pack_stream", "nybble_unpack_stream", "compress_bytestring",
             "decompress_bytestring", "nybble_compress", "nybble_decompress", "digit2int", "power", "array_max", "array_min", "dc_container_compress",
             "dc_container_decompress"]


@pytest.fixture(scope="module")
def built():
    import __graft_entry__ as g
    g.build()


def test_refapi_exports_reference_names(built):
    L = ctypes.CDLL(REFAPI)
    missing = [n for n in REF_NAMES if not hasattr(L, n)]
    assert not missing, missing
    assert os.access(os.path.join(HOST, "n_ary_huffman"), os.X_OK)
    assert os.access(os.path.join(HOST, "nybble_compression"), os.X_OK)


def test_digit2int_is_the_base64url_table(built):
    """digit2int() n_ary_huffman.c:428-455: the base64url alphabet, plus '+' and '/' (:441-445).  A host table look-up."""
    import base64
    L = ctypes.CDLL(REFAPI)
    L.digit2int.argtypes = [ctypes.c_char]
    L.digit2int.restype = ctypes.c_int
    alphabet = b"ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789-_"
    assert [L.digit2int(bytes([c])) for c in alphabet] == list(range(64))
    assert L.digit2int(b"+") == 62 and L.digit2int(b"/") == 63
    # the same table as RFC 4648: every 6-bit value through Python's encoder
    for v in range(64):
        ch = base64.urlsafe_b64encode(bytes([v << 2]))[:1]
        assert L.digit2int(ch) == v


def test_length_helpers_match_the_reference(built):
    """power / array_max / array_min (n_ary_huffman.c:1317-1379) against the unmodified reference functions, incl. the
    as-written quirk that the last slot is not looked at."""
    import numpy as np
    from oracle import pyoracle as O
    L = ctypes.CDLL(REFAPI)
    ip = ctypes.POINTER(ctypes.c_int)
    for f in (L.array_max, L.array_min):
        f.argtypes = [ctypes.c_int, ip]
        f.restype = ctypes.c_int
    L.power.argtypes = [ctypes.c_int, ctypes.c_int]
    L.power.restype = ctypes.c_int
    R = O.ref_huff() if O.have_ref() else None
    if R is not None:
        for f in (R.array_max, R.array_min):
            f.argtypes = [ctypes.c_int, ip]
            f.restype = ctypes.c_int
        R.power.argtypes = [ctypes.c_int, ctypes.c_int]
        R.power.restype = ctypes.c_int
    rng = np.random.default_rng(8)
    for _ in range(200):
        ln = rng.integers(0, 16, size=259).astype(np.int32) * (rng.random(259) < rng.random()).astype(np.int32)
        p = ln.ctypes.data_as(ip)
        body = ln[:258]
        assert L.array_max(258, p) == int(body.max())                                  # slot 258 is not looked at (:1336, :1361)
        assert L.array_min(258, p) == (int(body[body > 0].min()) if (body > 0).any() else 300)
        if R is not None:
            assert L.array_max(258, p) == R.array_max(258, p) and L.array_min(258, p) == R.array_min(258, p)
    for b, e in ((2, 0), (2, 10), (3, 5), (4, 15), (16, 7), (10, 9), (7, 1)):
        assert L.power(b, e) == b ** e
        if R is not None:
            assert L.power(b, e) == R.power(b, e)


def test_cli_aborts_without_device(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("checks the no-device behaviour")
    r = subprocess.run([os.path.join(HOST, "n_ary_huffman")], input=b"abc", capture_output=True)
    assert r.returncode != 0 and b"no CPU fallback" in r.stderr
    r = subprocess.run([os.path.join(HOST, "nybble_compression")], capture_output=True)
    assert r.returncode != 0 and b"Successful" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("radix,expect", [(None, 3), (2, 2), (4, 2), (16, 2), (5, 2)])
def test_n_ary_huffman_cli(built, radix, expect):
    from data_compression_b200 import synth
    thr, base = synth.zipf_7bit_spec()
    text = synth.host_stream(300000, 11, thr, base).tobytes()   # 7-bit, no NUL: what the reference accepts
    args = [os.path.join(HOST, "n_ary_huffman")] + (["--n", str(radix)] if radix else [])
    r = subprocess.run(args, input=text, capture_output=True, timeout=120)
    assert r.returncode == 0, r.stderr[-500:]
    out = r.stdout.decode(errors="replace")
    assert out.count("Successful test.") == expect
    if radix == 5:      # no payload form for this radix: tables on the GPU, raw block like the reference
        assert "pass-through raw data" in out.split("Starting next block")[1]
    else:               # the default is the reference's radix 3: 5 trits per byte
        assert "# compressed: 300000 ->" in out and "pass-through" not in out.split("Starting next block")[1]
    q = subprocess.run(args + ["--quiet"], input=text, capture_output=True, timeout=120)
    assert q.returncode == 0 and set(q.stdout.decode().split("\n")) <= {"Successful test.", ""}


@pytest.mark.gpu
def test_nybble_compression_cli(built):
    r = subprocess.run([os.path.join(HOST, "nybble_compression")], capture_output=True, timeout=60)
    assert r.returncode == 0 and r.stdout.decode().count("Successful test.") == 3   # nibbles, static table, adaptive contexts
    assert b"compressed 80 -> 57 bytes" in r.stdout           # the reference's own result on its fixed text
    data = np.random.default_rng(3).integers(0, 256, size=1 << 20, dtype=np.uint8).tobytes()
    r = subprocess.run([os.path.join(HOST, "nybble_compression"), "-"], input=data, capture_output=True, timeout=60)
    assert r.returncode == 0 and b"Successful test." in r.stdout


@pytest.mark.gpu
def test_write_nybble_verbatim_signature(built):
    """write_nybble(nybble, dest, offset) nybble_compression.c:1091-1114: offset 0 replaces the high nibble, 1 the low."""
    L = ctypes.CDLL(REFAPI)
    L.write_nybble.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_bool]
    L.write_nybble.restype = None
    b = ctypes.create_string_buffer(b"\x5a", 1)
    L.write_nybble(0xC, b, False)
    assert b.raw == b"\xca"
    L.write_nybble(0x3, b, True)
    assert b.raw == b"\xc3"


@pytest.mark.gpu
@pytest.mark.parametrize("radix", [2, 4, 16, 3, 10])
def test_container_blocks(built, radix):
    """The netstring container (n_ary_huffman.c:1705-1814 / :2014-2094): table block in the reference's own text form,
    data block, raw fall-back; written and read back through the library."""
    import numpy as np
    from data_compression_b200 import hostapi, synth
    L = ctypes.CDLL(REFAPI)
    L.dc_container_compress.restype = ctypes.c_size_t
    L.dc_container_compress.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t,
                                        ctypes.c_char_p, ctypes.c_size_t]
    L.dc_container_decompress.restype = ctypes.c_size_t
    L.dc_container_decompress.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_size_t, ctypes.c_char_p, ctypes.c_size_t]
    thr, base = synth.zipf_7bit_spec()
    text = synth.host_stream(200001, 21, thr, base).tobytes()
    h = hostapi.histogram(text)
    lengths = hostapi.huffman(h, radix).astype(np.int32)
    cap = len(text) + len(text) // 4 + 4096
    out = ctypes.create_string_buffer(cap + 1)
    n = L.dc_container_compress(radix, lengths.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), text, len(text), out, cap)
    blob = out.raw[:n]
    if radix == 10:     # no payload packing for this radix: the raw block, like the reference
        assert blob == b"%d:\n\n" % (len(text) + 2) + text + b","
    else:
        digits = "".join("0123456789ABCDEF"[int(x)] for x in lengths)
        assert blob.startswith(b"265:\nX258:" + digits.encode() + b",")   # the reference's table block (:1727-1747)
        assert n < len(text)
    back = ctypes.create_string_buffer(len(text) + 2)
    m = L.dc_container_decompress(radix, blob, n, back, len(text) + 2)
    assert m == len(text) and back.raw[:m] == text
    assert L.dc_container_decompress(radix, blob[:-3], n - 3, back, len(text) + 2) == ctypes.c_size_t(-1).value   # truncated
    both = blob + b"\n"   # the reader also takes the line feed the reference's reader asserts behind a block (:2049)
    assert L.dc_container_decompress(radix, both, len(both), back, len(text) + 2) == len(text)


@pytest.mark.gpu
def test_container_table_block_equals_the_reference_writer(built):
    """The table block 'X' byte for byte as the unmodified compress() formats it (n_ary_huffman.c:1705-1747): golden buffer
    from tests/golden/container.json -- behind a one-byte text the block is still in the reference's buffer from offset 7."""
    import json
    import numpy as np
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "container.json")))["cases"]
    L = ctypes.CDLL(REFAPI)
    L.dc_container_compress.restype = ctypes.c_size_t
    L.dc_container_compress.argtypes = [ctypes.c_int, ctypes.POINTER(ctypes.c_int), ctypes.c_char_p, ctypes.c_size_t,
                                        ctypes.c_char_p, ctypes.c_size_t]
    seen = 0
    for case in gold:
        text, n = bytes.fromhex(case["text"]), case["n"]
        if len(text) != 1 or n < 3:
            continue
        buf = bytes.fromhex(case["buffer"])
        lengths = np.array(case["lengths"], dtype=np.int32)
        ref_block = b"265:\nX2" + buf[7: 7 + 3 + 259 + 1]          # "58:" + one digit per length + ","
        assert ref_block.endswith(b",") and len(ref_block) == 270
        long_text = text * 20000                                     # same alphabet, long enough to be worth coding
        cap = len(long_text) + len(long_text) // 4 + 4096
        out = ctypes.create_string_buffer(cap + 1)
        m = L.dc_container_compress(n, lengths.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), long_text, len(long_text), out, cap)
        assert m != ctypes.c_size_t(-1).value and out.raw[:270] == ref_block, (n, out.raw[:40])
        seen += 1
    assert seen >= 2
