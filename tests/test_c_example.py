"""The plain-C program on the C ABI (examples/multicell_uplink.c: pthread workers, one engine each, batched transport-block
decode with device-resident soft buffers, vectors from the engine's own encoder) builds and decodes its own transmissions."""
import json
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "examples", "multicell_uplink")


def test_c_example_is_built():
    import __graft_entry__ as g
    g.build()
    assert os.path.exists(BIN)


@pytest.mark.gpu
@pytest.mark.parametrize("pinned", [1, 0])
def test_c_example_runs(pinned):
    if not os.path.exists(BIN):
        pytest.skip("examples/multicell_uplink not built")
    out = subprocess.run([BIN, "2", "8", "3", str(pinned)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    d = json.loads(out.stdout.strip().splitlines()[-1])
    assert d["transport_blocks"] == 2 * 8 * 3
    assert d["crc_ok_but_payload_differs"] == 0
    assert d["crc_ok"] >= d["transport_blocks"] * 0.8     # rate 0.87 at this noise level: most blocks pass, a few need HARQ
