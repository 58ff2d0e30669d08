"""Torch7 binary serialisation (torch.save / torch.load, torch7 File.lua + generic/Tensor.c,
generic/Storage.c) for the object kinds the reference's checkpoints hold: numbers, booleans, strings,
tables and Float/Double/Long tensors.

Why it exists: the reference persists its state with torch.save (utils.lua:73-80 `safe_save`,
main.lua:181) and reads it back with torch.load (main.lua:146-148; mainviz.lua:12-15 loads
`parameters/means`, `parameters/vars` and `opt`).  A user switching to libvbnn.so keeps those
files readable by the unmodified Lua scripts.  No Torch7 exists in this image, so the format below is
restated from torch7's sources (un-vendored, like the rest of Torch7 -- SURVEY.md 8c) and checked by
a byte-level known-answer test plus round trips (tests/test_checkpoint_cpu.py).

Binary layout (little endian, 64-bit longs -- torch's default on x86-64):
  object   := int32 type, payload
  type     0 nil | 1 number (float64) | 2 string (int32 len, bytes) | 3 table | 4 torch object | 5 boolean (int32)
  table    := int32 index, [if first occurrence] int32 npairs, (object key, object value) * npairs
  torch    := int32 index, [if first occurrence] string "V 1", string class name, class payload
  Tensor   := int32 ndim, int64 size[ndim], int64 stride[ndim], int64 storageOffset (1-based), object storage
  Storage  := int64 n, raw elements
"""
from __future__ import annotations

import io
import struct

import numpy as np

_TENSOR = {"torch.FloatTensor": ("torch.FloatStorage", np.float32), "torch.DoubleTensor": ("torch.DoubleStorage", np.float64),
           "torch.LongTensor": ("torch.LongStorage", np.int64), "torch.CudaTensor": ("torch.CudaStorage", np.float32)}
_STORAGE = {v[0]: v[1] for v in _TENSOR.values()}
_BY_DTYPE = {np.dtype(np.float32): "torch.FloatTensor", np.dtype(np.float64): "torch.DoubleTensor",
             np.dtype(np.int64): "torch.LongTensor"}


class _Writer:
    def __init__(self, f):
        self.f = f
        self.index = 0
        self.seen = {}

    def i32(self, v):
        self.f.write(struct.pack("<i", int(v)))

    def i64(self, v):
        self.f.write(struct.pack("<q", int(v)))

    def string(self, s):
        b = s.encode() if isinstance(s, str) else bytes(s)
        self.i32(len(b))
        self.f.write(b)

    def _new_index(self, obj):
        self.index += 1
        self.seen[id(obj)] = self.index
        return self.index

    def obj(self, o):
        if o is None:
            self.i32(0)
        elif isinstance(o, (bool, np.bool_)):
            self.i32(5); self.i32(1 if o else 0)
        elif isinstance(o, (int, float, np.integer, np.floating)):
            self.i32(1); self.f.write(struct.pack("<d", float(o)))
        elif isinstance(o, (str, bytes)):
            self.i32(2); self.string(o)
        elif isinstance(o, np.ndarray) or hasattr(o, "detach"):
            self.tensor(o)
        elif isinstance(o, dict):
            self.table(o, list(o.items()))
        elif isinstance(o, (list, tuple)):
            self.table(o, [(i + 1, v) for i, v in enumerate(o)])       # Lua arrays are 1-based
        else:
            raise TypeError(f"t7: cannot serialise {type(o).__name__}")

    def table(self, o, pairs):
        self.i32(3)
        if id(o) in self.seen:
            self.i32(self.seen[id(o)])
            return
        self.i32(self._new_index(o))
        self.i32(len(pairs))
        for k, v in pairs:
            self.obj(k)
            self.obj(v)

    def tensor(self, t):
        if hasattr(t, "detach"):
            t = t.detach().cpu().numpy()
        a = np.ascontiguousarray(t)
        if a.dtype not in _BY_DTYPE:
            a = a.astype(np.float32)
        cls = _BY_DTYPE[a.dtype]
        self.i32(4)
        self.i32(self._new_index(t))
        self.string("V 1")
        self.string(cls)
        self.i32(a.ndim)
        for s in a.shape:
            self.i64(s)
        for s in a.strides:
            self.i64(s // a.itemsize)
        self.i64(1)                                                    # storageOffset, 1-based
        # the storage object
        self.i32(4)
        self.index += 1
        self.i32(self.index)
        self.string("V 1")
        self.string(_TENSOR[cls][0])
        self.i64(a.size)
        self.f.write(a.tobytes())


class _Reader:
    def __init__(self, f):
        self.f = f
        self.memo = {}

    def _read(self, fmt):
        n = struct.calcsize(fmt)
        b = self.f.read(n)
        if len(b) != n:
            raise EOFError("t7: truncated file")
        return struct.unpack(fmt, b)[0]

    def i32(self):
        return self._read("<i")

    def i64(self):
        return self._read("<q")

    def string(self):
        n = self.i32()
        return self.f.read(n).decode()

    def obj(self):
        t = self.i32()
        if t == 0:
            return None
        if t == 1:
            v = self._read("<d")
            return int(v) if float(v).is_integer() and abs(v) < 2 ** 53 else v
        if t == 2:
            return self.string()
        if t == 5:
            return self.i32() != 0
        if t == 3:
            idx = self.i32()
            if idx in self.memo:
                return self.memo[idx]
            out = {}
            self.memo[idx] = out
            for _ in range(self.i32()):
                k = self.obj()
                out[k] = self.obj()
            n = len(out)
            if n and all(isinstance(k, int) for k in out) and sorted(out) == list(range(1, n + 1)):
                lst = [out[i] for i in range(1, n + 1)]                # a Lua array
                self.memo[idx] = lst
                return lst
            return out
        if t == 4:
            idx = self.i32()
            if idx in self.memo:
                return self.memo[idx]
            version = self.string()
            cls = self.string() if version.startswith("V ") else version
            if cls in _TENSOR:
                nd = self.i32()
                size = [self.i64() for _ in range(nd)]
                stride = [self.i64() for _ in range(nd)]
                off = self.i64() - 1
                st = self.obj()
                if st is None or nd == 0:
                    out = np.zeros(size, dtype=_TENSOR[cls][1])
                else:
                    out = np.lib.stride_tricks.as_strided(st[off:], shape=size, strides=[s * st.itemsize for s in stride]).copy()
                self.memo[idx] = out
                return out
            if cls in _STORAGE:
                n = self.i64()
                dt = np.dtype(_STORAGE[cls])
                out = np.frombuffer(self.f.read(n * dt.itemsize), dtype=dt).copy()
                self.memo[idx] = out
                return out
            raise TypeError(f"t7: unsupported torch class {cls}")
        raise TypeError(f"t7: unsupported type tag {t}")


def dumps(obj) -> bytes:
    b = io.BytesIO()
    _Writer(b).obj(obj)
    return b.getvalue()


def loads(data: bytes):
    return _Reader(io.BytesIO(data)).obj()


def save(filename, obj):
    """torch.save(filename, obj) (binary mode)."""
    with open(filename, "wb") as f:
        _Writer(f).obj(obj)


def load(filename):
    """torch.load(filename) for files written by save() or by Torch7 with the object kinds above."""
    with open(filename, "rb") as f:
        return _Reader(f).obj()
