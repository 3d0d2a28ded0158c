"""GPU: a few seconds of each randomised cross-check (tools/fuzz*.py: random alphabets, skews, radices, sizes, bit phases,
damaged streams, strings and batches) against the oracle, with fixed seeds."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("tool,env", [("fuzz.py", {}), ("fuzz_text.py", {}), ("fuzz_host.py", {"DC_PIPE_CHUNK_MIB": "1"})])
def test_fuzz_tool_runs_clean(tool, env):
    e = dict(os.environ, SEED="5", **env)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", tool), "6"], capture_output=True, text=True, env=e, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "mismatches 0" in r.stdout
