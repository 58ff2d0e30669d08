"""CPU tests of the interop formats: Torch7 binary serialisation of the
checkpoint files (utils.lua:73-80, mainviz.lua:12-15) and the logger's one-value-per-line metric
files (logger.lua:18-26, read by visualize.py:25-31)."""
import os
import struct

import numpy as np

from vbnn_b200 import checkpoint, logger, t7


def test_t7_float_tensor_known_answer_bytes():
    """Byte-level known answer built by hand from torch7's File.lua / generic/Tensor.c / Storage.c."""
    a = np.array([[1, 2], [3, 4]], dtype=np.float32)
    s = lambda x: struct.pack("<i", len(x)) + x
    expect = (struct.pack("<ii", 4, 1) + s(b"V 1") + s(b"torch.FloatTensor") +
              struct.pack("<i", 2) + struct.pack("<qq", 2, 2) + struct.pack("<qq", 2, 1) + struct.pack("<q", 1) +
              struct.pack("<ii", 4, 2) + s(b"V 1") + s(b"torch.FloatStorage") + struct.pack("<q", 4) +
              struct.pack("<4f", 1, 2, 3, 4))
    assert t7.dumps(a) == expect
    assert np.array_equal(t7.loads(expect), a)
    # scalars, strings, booleans, nil
    assert t7.dumps(3) == struct.pack("<i", 1) + struct.pack("<d", 3.0)
    assert t7.dumps("ab") == struct.pack("<ii", 2, 2) + b"ab"
    assert t7.dumps(True) == struct.pack("<ii", 5, 1)
    assert t7.dumps(None) == struct.pack("<i", 0)


def test_t7_round_trip_tables_and_tensors(tmp_path):
    rng = np.random.RandomState(0)
    obj = {"opt": {"B": 1e6, "hidden": [1200, 1200], "classes": list("0123456789"), "cuda": True,
                   "state": {"learningRate": 0.001}, "network_name": "exp"},
           "0.means": rng.randn(5, 7).astype(np.float32), "0.lvars": rng.randn(5, 7).astype(np.float32),
           "0.t": 3, "d": rng.randn(4).astype(np.float64), "idx": np.arange(6, dtype=np.int64).reshape(2, 3)}
    f = tmp_path / "model"
    t7.save(f, obj)
    back = t7.load(f)
    assert back["opt"]["hidden"] == [1200, 1200] and back["opt"]["classes"] == list("0123456789")
    assert back["opt"]["cuda"] is True and back["opt"]["state"]["learningRate"] == 0.001
    assert back["0.t"] == 3
    for k in ("0.means", "0.lvars", "d", "idx"):
        assert back[k].dtype == obj[k].dtype and np.array_equal(back[k], obj[k])


def test_t7_reads_strided_views_and_shared_storage():
    """A tensor written by Torch7 may be a view (non-contiguous strides, storage offset)."""
    s = lambda x: struct.pack("<i", len(x)) + x
    data = np.arange(12, dtype=np.float32)
    blob = (struct.pack("<ii", 4, 1) + s(b"V 1") + s(b"torch.FloatTensor") + struct.pack("<i", 2) +
            struct.pack("<qq", 3, 2) + struct.pack("<qq", 1, 4) + struct.pack("<q", 2) +       # 3x2, strides (1,4), offset 2
            struct.pack("<ii", 4, 2) + s(b"V 1") + s(b"torch.FloatStorage") + struct.pack("<q", 12) + data.tobytes())
    got = t7.loads(blob)
    want = np.array([[1, 5], [2, 6], [3, 7]], dtype=np.float32)
    assert np.array_equal(got, want)


def test_safe_save_rotates_old_file(tmp_path):
    d = str(tmp_path / "exp")
    checkpoint.safe_save({"a": 1}, d, "model")
    checkpoint.safe_save({"a": 2}, d, "model")                          # utils.lua:76-78: mv model model.old
    assert t7.load(os.path.join(d, "model"))["a"] == 2
    assert t7.load(os.path.join(d, "model.old"))["a"] == 1


def test_logger_writes_one_value_per_line(tmp_path):
    """logger.lua:18-26 + the reader visualize.py:25-31 uses."""
    Log = logger.init(str(tmp_path / "exp"))
    vals = [0.125, 3.0, 1e-9, -2.5e7, 1 / 3]
    for v in vals:
        Log.add("var hat", v)
        Log.add("devacc", 2 * v)
    Log.flush()
    lines = open(tmp_path / "exp" / "var hat").read().splitlines()
    assert lines[:2] == ["0.125", "3"] and len(lines) == 5              # Lua tostring(): "%.14g"
    got = logger.read_data(str(tmp_path / "exp" / "var hat"))
    assert np.allclose(got, vals, rtol=1e-13)
    Log.close()
    # a fresh non-append logger truncates (logger.lua:20-22); append mode keeps (main.lua:148)
    Log = logger.init(str(tmp_path / "exp"), append=True)
    Log.add("var hat", 7)
    Log.close()
    assert len(logger.read_data(str(tmp_path / "exp" / "var hat"))) == 6
    Log = logger.init(str(tmp_path / "exp"))
    Log.add("var hat", 8)
    Log.close()
    assert logger.read_data(str(tmp_path / "exp" / "var hat")) == [8.0]
    logger.Log = None


def test_snr_pruning_restatement():
    """mainviz.lua:20-24 on known values."""
    means = np.array([0.0, 1e-4, 0.5, -1e-5], dtype=np.float32)
    vars_ = np.array([1e-2, 1e-2, 1e-2, 1e-6], dtype=np.float32)
    mask, count, mean_v, mean_pv = checkpoint.snr_pruned(means, vars_, 0.005)
    assert mask.tolist() == [1.0, 1.0, 0.0, 0.0] and count == 2.0
    assert abs(mean_pv - 0.005) < 1e-9
