import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with open(os.path.join(GOLDEN, name)) as f:
        return json.load(f)


def dense(sparse: dict, n: int = 259, dtype=np.int64) -> np.ndarray:
    a = np.zeros(n, dtype=dtype)
    for k, v in sparse.items():
        a[int(k)] = v
    return a


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def dc():
    """The product package with its CUDA library loaded; GPU tests fail (not skip) if it cannot run."""
    import torch
    import data_compression_b200 as dc
    dc.lib()
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    torch.cuda.set_device(0)
    return dc


@pytest.fixture(scope="session")
def table_cases():
    return load_golden("huffman_tables.json")["cases"]
