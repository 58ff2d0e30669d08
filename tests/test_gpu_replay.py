"""The layer-level C ABI driven by a NON-PYTHON host (tools/replay.cu -> tools/replay, links libvbnn.so + cudart
only) in the order the reference's Lua issues the calls (mlp.lua:62-84,117-142; main.lua:28-40), with the
reference's other modules (ReLU, output nn.Linear, LogSoftMax, criterion, optim.sgd, gradParameters:zero()) as
foreign kernels on the legacy default stream in between, and the VB layers bound into a getParameters()-style
flat storage (vbnn_layer_bind).  Pass = the oracle fixture tests/golden/mlp_weight.npz reproduced."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tools", "replay")
FIX = os.path.join(ROOT, "tests", "golden", "mlp_weight.replay.bin")


def read_replay_bin(path):
    out = {}
    with open(path, "rb") as f:
        assert f.read(4) == b"VBRP"
        (count,) = struct.unpack("<I", f.read(4))
        for _ in range(count):
            (nl,) = struct.unpack("<I", f.read(4))
            name = f.read(nl).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            dims = struct.unpack(f"<{nd}I", f.read(4 * nd))
            n = int(np.prod(dims))
            out[name] = np.frombuffer(f.read(4 * n), dtype="<f4").reshape(dims)
    return out


def test_replay_fixture_is_the_npz_fixture():
    """(CPU) the flat file the C++ host reads holds exactly the committed oracle fixture."""
    g = np.load(os.path.join(ROOT, "tests", "golden", "mlp_weight.npz"))
    b = read_replay_bin(FIX)
    assert sorted(b) == sorted(g.files)
    for k in g.files:
        assert np.array_equal(b[k].ravel(), np.atleast_1d(g[k]).astype(np.float32).ravel()), k


def test_replay_host_is_built_and_links_only_the_c_abi():
    """(CPU) built by __graft_entry__.build(); depends on libvbnn.so + the CUDA runtime, nothing of Python / torch."""
    assert os.path.isfile(BIN), "tools/replay missing: run __graft_entry__.build()"
    r = subprocess.run(["ldd", BIN], capture_output=True, text=True)
    assert "libvbnn.so" in r.stdout
    assert "libtorch" not in r.stdout and "libpython" not in r.stdout


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["legacy", "blocking"])
def test_replay_matches_oracle_fixture(mode):
    r = subprocess.run([BIN, FIX, mode], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"REPLAY OK (stream mode: {mode})" in r.stdout
