"""The __host__ __device__ format logic (phyngsc_b200/csrc/phy_core.cuh) compiled for the CPU by
tests/mirror/mirror.cpp and checked against the oracle: tokeniser, classification, Huffman build and
serialisation, header layout, stream walkers and bit sinks.  The CUDA kernels call the same functions."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from phyngsc_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "mirror", "mirror.cpp")
LIB = os.path.join(HERE, "mirror", "libmirror.so")
CORE = os.path.join(HERE, "..", "phyngsc_b200", "csrc", "phy_core.cuh")
FAST = os.path.join(HERE, "..", "phyngsc_b200", "csrc", "phy_fast.cuh")


@pytest.fixture(scope="module")
def mirror():
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(CORE), os.path.getmtime(FAST)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-x", "c++", "-fPIC", "-shared", "-o", LIB, SRC])
    L = C.CDLL(LIB)
    L.mirror_compress_window.restype = C.c_int
    L.mirror_compress_window.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_int32, C.c_uint32, C.c_void_p,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p]
    L.mirror_compress_window_fast.restype = C.c_int
    L.mirror_compress_window_fast.argtypes = L.mirror_compress_window.argtypes + [C.c_uint32]
    L.mirror_huffman.restype = C.c_uint32
    L.mirror_huffman.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
    return L


@pytest.mark.parametrize("shape", ["36bp", "100bp", "100bp_huffdna", "150bp_paired", "var50_205", "var50_250", "title_stress",
                                   "degrade", "mixed_amb"])
@pytest.mark.parametrize("nbytes,rsize", [(300_000, 200_000), (2_500_000, 1 << 21)])
def test_core_logic_reproduces_oracle_window(shape, nbytes, rsize, mirror, oracle):
    d = synth.fastq(shape, 3, target_bytes=nbytes)
    rsize = min(rsize, d.size)
    want = oracle.compress_window(d, r_buffer_size=rsize)
    out = np.zeros(d.size + 65536, np.uint8); sec = np.zeros(4, np.uint32); nr = C.c_uint32(); bc = C.c_uint64()
    rc = mirror.mirror_compress_window(d.ctypes.data, d.size, rsize, 0, 500, 100000, out.ctypes.data, out.size, sec.ctypes.data,
                                       C.byref(nr), C.byref(bc))
    assert rc == 0
    assert sec.tolist() == [len(s) for s in want["sections"]]
    assert out[: int(sec.sum())].tobytes() == want["payload"]
    assert (nr.value, bc.value) == (want["n_records"], want["bytes_consumed"])


@pytest.mark.parametrize("shape", ["36bp", "100bp", "100bp_huffdna", "150bp_paired", "var50_205", "var50_250", "title_stress",
                                   "degrade", "mixed_amb"])
@pytest.mark.parametrize("lanes_per_record", [1, 2, 4, 8])
def test_single_walk_encoder_logic_reproduces_oracle_window(shape, lanes_per_record, mirror, oracle):
    """phy_fast.cuh (lane-private sinks, warp concatenation with carry, task slots, placement) emulated lane by lane, with
    1..8 lanes sharing a record's quality / DNA codes, against the oracle on a window of several tasks."""
    d = synth.fastq(shape, 5, target_bytes=1_200_000)
    rsize = min(1 << 20, d.size)
    want = oracle.compress_window(d, r_buffer_size=rsize)
    out = np.zeros(d.size + 65536, np.uint8); sec = np.zeros(4, np.uint32); nr = C.c_uint32(); bc = C.c_uint64()
    rc = mirror.mirror_compress_window_fast(d.ctypes.data, d.size, rsize, 0, 500, 100000, out.ctypes.data, out.size, sec.ctypes.data,
                                            C.byref(nr), C.byref(bc), lanes_per_record)
    if rc == 1000:
        pytest.skip("title bound beyond the lane-private staging: the GPU takes the two-walk kernels here")
    assert rc == 0
    assert sec.tolist() == [len(s) for s in want["sections"]]
    assert out[: int(sec.sum())].tobytes() == want["payload"]


def test_core_huffman_matches_oracle(mirror, oracle):
    rng = np.random.default_rng(11)
    for it in range(300):
        n = int(rng.choice([1, 2, 3, 5, 8, 41, 64, 100, 256, 300, 512]))
        f = (rng.integers(0, 3, n) * rng.integers(0, 1000, n)).astype(np.uint32) if it % 3 else rng.integers(0, 6, n).astype(np.uint32)
        try:
            code, ln, tree = oracle.huffman(f, True)
        except ValueError:
            continue
        cl = np.zeros(n, np.uint64); tb = np.zeros(4096, np.uint8)
        k = mirror.mirror_huffman(f.ctypes.data, n, cl.ctypes.data, tb.ctypes.data)
        assert tb[:k].tobytes() == tree
        assert ((cl >> np.uint64(32)).astype(np.uint32) == ln).all() and ((cl & np.uint64(0xFFFFFFFF)).astype(np.uint32) == code).all()
