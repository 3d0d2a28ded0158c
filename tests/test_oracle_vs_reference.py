"""CPU: differential test of the restatement against the UNMODIFIED reference functions compiled by
oracle/Makefile (oracle/_ref).  Skipped only if the prebuilt harness is absent (it is built wherever
/root/reference exists and travels to the GPU box with the snapshot)."""
import numpy as np
import pytest

RADICES = (2, 3, 4, 5, 10, 16)


@pytest.fixture(scope="module")
def ref(oracle):
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built (needs /root/reference at build time)")
    return oracle


def _random_hist(rng, t):
    h = np.zeros(259, dtype=np.int64)
    k = int(rng.integers(1, 257))
    idx = rng.choice(np.arange(0, 256), size=k, replace=False)
    mode = t % 5
    if mode == 0:
        h[idx] = rng.integers(1, 4, size=k)
    elif mode == 1:
        h[idx] = rng.integers(1, 1 << 20, size=k)
    elif mode == 2:
        h[idx] = (1e7 / (np.arange(1, k + 1) ** 1.1)).astype(np.int64) + 1
    elif mode == 3:
        h[idx] = 1
    else:
        h[idx] = 2 ** rng.integers(0, 12, size=k)
    return h


def test_tables_differential(ref):
    rng = np.random.default_rng(20261018)
    for t in range(120):
        h = _random_hist(rng, t)
        for n in RADICES:
            a = ref.huffman(h.astype(np.uint64), n)
            b = ref.ref_huffman(h.astype(np.int32), n)
            assert np.array_equal(a, b), (t, n)
            if 0 < a.max() < 16 and n ** int(a.max()) < 2 ** 31:
                el, ev, st = ref.convert_lengths_to_encode_table(a, n)
                el2, ev2 = ref.ref_convert_lengths_to_encode_table(b, n)
                assert st == 0 and np.array_equal(el, el2) and np.array_equal(ev, ev2), (t, n)


def test_histogram_differential(ref):
    rng = np.random.default_rng(5)
    text = bytes(rng.integers(1, 127, size=20000, dtype=np.uint8).tolist())
    assert np.array_equal(ref.ref_histogram(text).astype(np.uint64), ref.histogram_cstr(text))


def test_reference_selftests_run_to_completion(ref):
    # with -DNDEBUG the reference's own main() paths finish; stdin is empty here
    assert ref.ref_nybble().ref_nybble_selftest() == 0


def test_nybble_differential(ref):
    rng = np.random.default_rng(11)
    alphabet = np.frombuffer(b" etaoinsxyzQ.,ABC", dtype=np.uint8)
    for _ in range(60):
        k = int(rng.integers(1, 400))
        text = bytes(rng.choice(alphabet, size=k).tolist())
        comp = ref.ref_compress_bytestring(text, False)
        assert ref.nybble_static_compress(text) == comp
        assert ref.nybble_static_decompress(comp) == text
        assert ref.ref_decompress_bytestring(comp, False) == text
        # modify == true: nybble_compress() / nybble_decompress() (:1134, :1117), 16 move-to-front contexts
        comp = ref.ref_compress_bytestring(text, True)
        assert ref.nybble_adaptive_compress(text) == comp
        assert ref.nybble_adaptive_decompress(comp) == text
        assert ref.ref_decompress_bytestring(comp, True) == text
    every = bytes(rng.permutation(np.arange(1, 128, dtype=np.uint8)).tolist()) * 3   # every context, every 7-bit byte
    comp = ref.ref_compress_bytestring(every, True)
    assert ref.nybble_adaptive_compress(every) == comp and ref.nybble_adaptive_decompress(comp) == every
    s = rng.integers(0, 16, size=257).astype(np.uint8)
    assert np.array_equal(ref.ref_write_nybble_stream(s), ref.nybble_pack(s))
