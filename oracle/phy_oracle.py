"""ctypes binding for oracle/liboracle.so (phy_oracle.c) and oracle/_ref (the compiled reference).

TEST INFRASTRUCTURE: see oracle/phy_oracle.h for what each entry point restates.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
REF_BIN = os.path.join(_HERE, "_ref", "phyNGSC_ref")
REF_KAT = os.path.join(_HERE, "_ref", "libphyref_kat.so")

WINDOW_BYTES = 1 << 23
BLOCK_BYTES = 1 << 23
RECORD_CAP = 100000


def build(force=False):
    """Compile liboracle.so (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(os.path.join(_HERE, "phy_oracle.c")):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    if os.path.exists("/root/reference/phyNGSC.cpp"):
        subprocess.check_call(["make", "-C", _HERE, "ref"], stdout=subprocess.DEVNULL)


class _Subblock(C.Structure):
    _fields_ = [("n_records", C.c_uint32), ("bytes_consumed", C.c_uint64), ("len", C.c_uint32 * 4),
                ("payload", C.POINTER(C.c_uint8)), ("payload_len", C.c_uint32), ("warnings", C.c_uint32)]


class _Rank(C.Structure):
    _fields_ = [("n_subblocks", C.c_uint32), ("sb_bytes", C.POINTER(C.c_uint8)), ("sb_off", C.POINTER(C.c_uint64)),
                ("sb_records", C.POINTER(C.c_uint32)), ("sb_win_off", C.POINTER(C.c_uint64)),
                ("sb_win_len", C.POINTER(C.c_uint64)), ("sb_rec_start", C.POINTER(C.c_uint32)),
                ("sb_overlap", C.POINTER(C.c_int32)), ("sb_len", C.POINTER(C.c_uint32 * 4)),
                ("n_blocks", C.c_uint32), ("blk_bytes", C.POINTER(C.c_uint8)), ("blk_off", C.POINTER(C.c_uint64)),
                ("last_block_size", C.c_uint32), ("wr_overlap", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.phy_oracle_huffman.restype = C.c_uint32
        L.phy_oracle_huffman.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32]
        L.phy_oracle_compress_window.restype = C.c_int
        L.phy_oracle_compress_window.argtypes = [C.c_void_p, C.c_uint64, C.c_int64, C.c_uint32, C.c_int32, C.c_uint32,
                                                 C.POINTER(_Subblock)]
        L.phy_oracle_subblock_free.argtypes = [C.POINTER(_Subblock)]
        L.phy_oracle_compress_rank.restype = C.c_int
        L.phy_oracle_compress_rank.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_uint64, C.c_uint64,
                                               C.c_uint32, C.POINTER(_Rank)]
        L.phy_oracle_rank_free.argtypes = [C.POINTER(_Rank)]
        L.phy_oracle_make_footer.restype = C.c_int32
        L.phy_oracle_make_footer.argtypes = [C.c_int, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                             C.c_void_p, C.c_void_p, C.c_uint32]
        L.phy_oracle_make_header.restype = C.c_uint32
        L.phy_oracle_make_header.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_uint32,
                                             C.c_void_p, C.c_uint32]
        _lib = L
    return _lib


def huffman(freq, compact=True):
    """-> (codes, lens, serialised tree bytes) as HuffmanEncoder::Complete + StoreTree produce them."""
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    n = f.size
    code = np.zeros(n, np.uint32)
    ln = np.zeros(n, np.uint32)
    tree = np.zeros(16 + 2 * n * 2 + 64, np.uint8)
    k = lib().phy_oracle_huffman(f.ctypes.data, n, int(compact), code.ctypes.data, ln.ctypes.data, tree.ctypes.data, tree.size)
    if k == 0:
        raise ValueError("oracle huffman failed (n out of range or code longer than 32 bits)")
    return code, ln, tree[:k].tobytes()


def compress_window(buf, r_buffer_size=None, rec_start=0, overlap=500, record_cap=RECORD_CAP):
    """One subblock.  -> dict(n_records, bytes_consumed, sections=[info,title,quality,dna], payload)."""
    a = np.frombuffer(buf, dtype=np.uint8) if not isinstance(buf, np.ndarray) else buf
    a = np.ascontiguousarray(a)
    if r_buffer_size is None:
        r_buffer_size = a.size
    sb = _Subblock()
    rc = lib().phy_oracle_compress_window(a.ctypes.data, a.size, r_buffer_size, rec_start, overlap, record_cap, C.byref(sb))
    if rc:
        raise OracleError(rc)
    payload = bytes(bytearray(sb.payload[: sb.payload_len]))
    lens = list(sb.len)
    secs, o = [], 0
    for ln in lens:
        secs.append(payload[o:o + ln]); o += ln
    out = dict(n_records=sb.n_records, bytes_consumed=sb.bytes_consumed, sections=secs, payload=payload, warnings=sb.warnings)
    lib().phy_oracle_subblock_free(C.byref(sb))
    return out


class OracleError(Exception):
    def __init__(self, rc):
        super().__init__(f"oracle error {rc}")
        self.rc = rc


def compress_rank(file_bytes, np_ranks, rank, window_bytes=WINDOW_BYTES, block_bytes=BLOCK_BYTES, record_cap=RECORD_CAP, threads=1):
    """Everything one rank does.  -> dict(subblocks=[bytes], blocks=[bytes], records=[...], windows=[(off,len,rec_start,overlap)],
    last_block_size, wr_overlap, section_lens)."""
    a = np.frombuffer(file_bytes, dtype=np.uint8) if not isinstance(file_bytes, np.ndarray) else file_bytes
    a = np.ascontiguousarray(a)
    r = _Rank()
    lib().phy_oracle_set_threads(int(threads))
    try:
        rc = lib().phy_oracle_compress_rank(a.ctypes.data, a.size, np_ranks, rank, window_bytes, block_bytes, record_cap, C.byref(r))
    finally:
        lib().phy_oracle_set_threads(1)
    if rc:
        raise OracleError(rc)
    ns, nb = r.n_subblocks, r.n_blocks
    sb_off = [r.sb_off[i] for i in range(ns + 1)]
    blk_off = [r.blk_off[i] for i in range(nb + 1)]
    sbb = np.ctypeslib.as_array(r.sb_bytes, shape=(max(sb_off[-1], 1),))[: sb_off[-1]].tobytes() if ns else b""
    bb = np.ctypeslib.as_array(r.blk_bytes, shape=(max(blk_off[-1], 1),))[: blk_off[-1]].tobytes() if nb else b""
    out = dict(
        subblocks=[sbb[sb_off[i]:sb_off[i + 1]] for i in range(ns)],
        blocks=[bb[blk_off[i]:blk_off[i + 1]] for i in range(nb)],
        records=[r.sb_records[i] for i in range(ns)],
        windows=[(r.sb_win_off[i], r.sb_win_len[i], r.sb_rec_start[i], r.sb_overlap[i]) for i in range(ns)],
        section_lens=[list(r.sb_len[i]) for i in range(ns)],
        last_block_size=r.last_block_size, wr_overlap=r.wr_overlap)
    lib().phy_oracle_rank_free(C.byref(r))
    return out


def make_footer(np_ranks, fastq_size, n_blocks, n_subblocks, overlaps, block_order, lb_sizes):
    ov = np.ascontiguousarray(overlaps, np.int32); bo = np.ascontiguousarray(block_order, np.int32)
    lb = np.ascontiguousarray(lb_sizes, np.uint32)
    out = np.zeros(64 + 4 * (len(bo) + 2 * np_ranks), np.uint8)
    k = lib().phy_oracle_make_footer(np_ranks, fastq_size, n_blocks, n_subblocks, ov.ctypes.data, bo.ctypes.data, lb.ctypes.data,
                                     out.ctypes.data, out.size)
    if k < 0:
        raise OracleError(k)
    return out[:k].tobytes()


# ---- the compiled reference ------------------------------------------------------------------
def have_reference():
    return os.path.exists(REF_BIN)


def run_reference(fastq_path, ngsc_path, np_ranks=2, threads=1, timeout=600):
    """Run the UNMODIFIED reference driver (oracle/_ref/phyNGSC_ref) under the fork-based MPI stand-in."""
    if os.path.exists(ngsc_path):
        os.remove(ngsc_path)  # Q15: the reference does not truncate
    env = dict(os.environ, PHY_SHIM_NP=str(np_ranks), OMP_NUM_THREADS=str(threads))
    p = subprocess.run([REF_BIN, fastq_path, ngsc_path, str(threads)], env=env, capture_output=True, text=True, timeout=timeout)
    if p.returncode != 0:
        raise RuntimeError(f"reference exited {p.returncode}: {p.stdout[-400:]} {p.stderr[-400:]}")
    return p.stdout


_kat = None


def ref_kat():
    global _kat
    if _kat is None:
        K = C.CDLL(REF_KAT)
        K.ref_huffman.restype = C.c_int
        K.ref_huffman.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_void_p]
        K.ref_bitstream.restype = C.c_int
        K.ref_bitstream.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p]
        if hasattr(K, "ref_decode_subblock"):
            K.ref_decode_subblock.restype = C.c_longlong
            K.ref_decode_subblock.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_ulonglong]
        _kat = K
    return _kat


def ref_decode_subblock(payload, cap):
    """FASTQ text of one subblock payload, decoded by the reference's own Fetch* functions (oracle/ref_kat.cpp)."""
    p = np.ascontiguousarray(np.frombuffer(payload, np.uint8) if not isinstance(payload, np.ndarray) else payload)
    out = np.empty(int(cap), np.uint8)
    n = ref_kat().ref_decode_subblock(p.ctypes.data, p.size, out.ctypes.data, out.size)
    if n < 0:
        raise RuntimeError(f"ref_decode_subblock rc={n}")
    return out[:n]


def ref_huffman(freq, compact=True):
    f = np.ascontiguousarray(freq, dtype=np.uint32)
    n = f.size
    code = np.zeros(n, np.uint32); ln = np.zeros(n, np.uint32)
    tree = np.zeros(16 + 4 * n + 64, np.uint8); tl = C.c_uint32()
    rc = ref_kat().ref_huffman(f.ctypes.data, n, int(compact), code.ctypes.data, ln.ctypes.data, tree.ctypes.data, tree.size, C.byref(tl))
    if rc:
        raise RuntimeError(f"ref_huffman rc={rc}")
    return code, ln, tree[: tl.value].tobytes()
