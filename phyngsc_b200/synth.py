"""Synthetic FASTQ of the shapes named by BASELINE.json / SURVEY.md 8(d) (csrc/fastq_synth.c)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "csrc", "fastq_synth.c")
_LIB = os.path.join(_HERE, "csrc", "libphysynth.so")

SHAPES = {"36bp": 1, "100bp": 3, "100bp_huffdna": 30, "150bp_paired": 4, "var50_250": 5, "var50_205": 50,
          "title_stress": 60, "degrade": 61, "mixed_amb": 62}

_lib = None


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        cc = "/usr/bin/gcc" if os.path.exists("/usr/bin/gcc") else "gcc"
        subprocess.check_call([cc, "-O2", "-fPIC", "-shared", "-o", _LIB, _SRC])


def _get():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.phy_synth_fastq.restype = C.c_uint64
        L.phy_synth_fastq.argtypes = [C.c_int, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        _lib = L
    return _lib


def fastq(shape, seed, target_bytes=0, max_records=0, out=None):
    """Generate whole records until target_bytes (or max_records) is reached.  Returns a uint8 array
    (a view into `out` when given).  `shape` is a key of SHAPES or its integer id."""
    sid = SHAPES[shape] if isinstance(shape, str) else int(shape)
    if target_bytes == 0:
        target_bytes = max_records * 2048
    cap = target_bytes + 4096
    buf = np.empty(cap, np.uint8) if out is None else out
    n = C.c_uint64()
    k = _get().phy_synth_fastq(sid, seed, target_bytes, max_records, buf.ctypes.data, min(cap, buf.size), C.byref(n))
    if k == 0:
        raise ValueError(f"unknown shape {shape!r} or buffer too small")
    return buf[:k]
