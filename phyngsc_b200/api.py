"""ctypes binding of include/phyngsc_b200.h -- the host-side mirror of the seam cut into the reference's
subblock loop (phyNGSC.cpp:168-840).  Everything computes on the GPU through the C ABI; importing works
without a GPU, creating a Context does not."""
import ctypes as C
import os

import numpy as np

from . import build as _build

WINDOW_BYTES = 1 << 23   # READ_BUFFER_SIZE, defs.h:20
BLOCK_BYTES = 1 << 23    # WRITE_BUFFER_SIZE, defs.h:21
OVERLAP = 500            # phyNGSC.cpp:48
RECORD_CAP = 100000      # phyNGSC.cpp:51 at threads = 1


class RegionParams(C.Structure):
    _fields_ = [("file_size", C.c_uint64), ("np", C.c_int32), ("rank", C.c_int32), ("window_bytes", C.c_uint64),
                ("overlap", C.c_uint32), ("record_cap", C.c_uint32), ("threads", C.c_uint32), ("reserved", C.c_uint32)]


class SubblockDesc(C.Structure):
    _fields_ = [("win_off", C.c_uint64), ("win_len", C.c_uint64), ("rec_start", C.c_uint32), ("overlap", C.c_int32),
                ("n_records", C.c_uint32), ("warnings", C.c_uint32), ("bytes_consumed", C.c_uint64), ("sec_len", C.c_uint32 * 4),
                ("out_off", C.c_uint64), ("out_len", C.c_uint32), ("status", C.c_int32)]


class RegionResult(C.Structure):
    _fields_ = [("n_subblocks", C.c_uint32), ("n_batches", C.c_uint32), ("bytes_in", C.c_uint64), ("bytes_out", C.c_uint64),
                ("out_used", C.c_uint64), ("wr_overlap", C.c_int32), ("kernel_launches", C.c_uint32), ("kernel_ms", C.c_float),
                ("h2d_ms", C.c_float), ("d2h_ms", C.c_float)]


EXPORTS = ["phy_device_count", "phy_ctx_create", "phy_ctx_destroy", "phy_compress_region", "phy_upload", "phy_compress_resident", "phy_download",
           "phy_find_first_record", "phy_device_input", "phy_device_output", "phy_host_alloc", "phy_host_free",
           "phy_make_block_header", "phy_make_footer", "phy_debug_read", "phy_profile", "phy_profile_read", "phy_strerror", "phy_last_error", "phy_abi_version", "phy_decode_subblock", "phy_compress_region_streamed", "phy_compress_stream", "phy_stream_prepare"]

_lib = None


class PhyError(RuntimeError):
    def __init__(self, code, detail=""):
        super().__init__(f"phyngsc_b200 error {code}: {detail}")
        self.code = code


def lib():
    """Loads csrc/libphyngsc_b200.so (building it first if the sources are newer).  No fallback: a missing
    or unloadable library is an error."""
    global _lib
    if _lib is None:
        path = _build.build_lib()
        L = C.CDLL(path)
        L.phy_ctx_create.restype = C.c_int
        L.phy_ctx_create.argtypes = [C.POINTER(C.c_void_p), C.c_int, C.c_uint64, C.c_uint32]
        L.phy_ctx_destroy.argtypes = [C.c_void_p]
        L.phy_device_count.restype = C.c_int
        L.phy_compress_region.restype = C.c_int
        L.phy_compress_region.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(RegionParams), C.c_void_p, C.c_uint64,
                                          C.POINTER(SubblockDesc), C.POINTER(C.c_uint32), C.POINTER(RegionResult)]
        L.phy_upload.restype = C.c_int
        L.phy_upload.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.phy_compress_resident.restype = C.c_int
        L.phy_compress_resident.argtypes = [C.c_void_p, C.c_uint64, C.POINTER(RegionParams), C.POINTER(SubblockDesc),
                                            C.POINTER(C.c_uint32), C.POINTER(RegionResult)]
        L.phy_download.restype = C.c_int
        L.phy_download.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.phy_find_first_record.restype = C.c_int64
        L.phy_find_first_record.argtypes = [C.c_void_p, C.c_uint64]
        L.phy_device_input.restype = C.c_void_p
        L.phy_device_input.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.phy_device_output.restype = C.c_void_p
        L.phy_device_output.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.phy_host_alloc.restype = C.c_void_p
        L.phy_host_alloc.argtypes = [C.c_uint64]
        L.phy_host_free.argtypes = [C.c_void_p]
        L.phy_make_block_header.restype = C.c_uint32
        L.phy_make_block_header.argtypes = [C.c_int32] * 5 + [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32]
        L.phy_make_footer.restype = C.c_int32
        L.phy_make_footer.argtypes = [C.c_int32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                      C.c_uint32]
        L.phy_debug_read.restype = C.c_int64
        L.phy_debug_read.argtypes = [C.c_void_p, C.c_char_p, C.c_uint64, C.c_void_p, C.c_uint64]
        L.phy_profile.restype = C.c_int
        L.phy_profile.argtypes = [C.c_void_p, C.c_int]
        L.phy_profile_read.restype = C.c_int
        L.phy_profile_read.argtypes = [C.c_void_p, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
        L.phy_strerror.restype = C.c_char_p
        L.phy_strerror.argtypes = [C.c_int]
        L.phy_last_error.restype = C.c_char_p
        L.phy_last_error.argtypes = [C.c_void_p]
        L.phy_abi_version.restype = C.c_int
        L.phy_decode_subblock.restype = C.c_int64
        L.phy_decode_subblock.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64]
        _lib = L
    return _lib


def pinned_array(nbytes):
    """uint8 numpy array over page-locked host memory (phy_host_alloc).  Keep the returned object alive."""
    p = lib().phy_host_alloc(nbytes)
    if not p:
        raise MemoryError("phy_host_alloc failed")
    buf = (C.c_uint8 * nbytes).from_address(p)
    arr = np.frombuffer(buf, dtype=np.uint8)
    arr_holder = _Pinned(p, arr)
    return arr_holder


class _Pinned:
    def __init__(self, ptr, arr):
        self.ptr, self.array = ptr, arr

    def free(self):
        if self.ptr:
            self.array = None
            lib().phy_host_free(self.ptr)
            self.ptr = None


def region_params(file_size, np_ranks, rank, window_bytes=WINDOW_BYTES, overlap=OVERLAP, record_cap=RECORD_CAP, threads=1):
    return RegionParams(file_size, np_ranks, rank, window_bytes, overlap, record_cap, threads, 0)


def region_slice(file_size, np_ranks, rank, slack=0, overlap=OVERLAP):
    """Byte range [start, end) of the file that rank's region needs (phyNGSC.cpp:113-124) plus `slack`
    read-ahead bytes, clipped at EOF."""
    region = file_size // np_ranks
    start = rank * region
    end = file_size if rank == np_ranks - 1 else min(file_size, start + region + overlap + slack)
    return start, end


class Context:
    """One GPU context (phy_ctx).  compress_region() is the call a host driver makes per rank."""

    def __init__(self, device=0, max_batch_bytes=0, max_subblocks=0):
        self._h = C.c_void_p()
        rc = lib().phy_ctx_create(C.byref(self._h), device, max_batch_bytes, max_subblocks)
        if rc:
            raise PhyError(rc, lib().phy_strerror(rc).decode())

    def close(self):
        if self._h:
            lib().phy_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self, rc):
        return PhyError(rc, (lib().phy_last_error(self._h) or b"").decode() or lib().phy_strerror(rc).decode())

    def compress_region(self, region, params, out=None, max_descs=4096, check=True):
        """region: uint8 array starting at the rank's p_wr_start.  Returns (descs, out, result); payload i is
        out[d.out_off : d.out_off + d.out_len]."""
        region = np.ascontiguousarray(region, dtype=np.uint8)
        if out is None:
            out = np.empty(region.size // 2 + (1 << 20), np.uint8)
        descs = (SubblockDesc * max_descs)()
        n = C.c_uint32(max_descs)
        res = RegionResult()
        rc = lib().phy_compress_region(self._h, region.ctypes.data, region.size, C.byref(params), out.ctypes.data, out.size, descs,
                                       C.byref(n), C.byref(res))
        if rc and check:
            raise self._err(rc)
        return list(descs[: n.value]), out, res

    def upload(self, region):
        region = np.ascontiguousarray(region, dtype=np.uint8)
        rc = lib().phy_upload(self._h, region.ctypes.data, region.size)
        if rc:
            raise self._err(rc)

    def compress_resident(self, region_len, params, max_descs=4096):
        descs = (SubblockDesc * max_descs)()
        n = C.c_uint32(max_descs)
        res = RegionResult()
        rc = lib().phy_compress_resident(self._h, region_len, C.byref(params), descs, C.byref(n), C.byref(res))
        if rc:
            raise self._err(rc)
        return list(descs[: n.value]), res

    def download(self, out):
        n = C.c_uint64()
        rc = lib().phy_download(self._h, out.ctypes.data, out.size, C.byref(n))
        if rc:
            raise self._err(rc)
        return n.value

    def profile(self, enable=True):
        rc = lib().phy_profile(self._h, int(enable))
        if rc:
            raise self._err(rc)

    def profile_read(self):
        """-> {stage: mean ms per batch} since profile(True)."""
        names = (C.c_char_p * 32)()
        ms = (C.c_float * 32)()
        n = lib().phy_profile_read(self._h, names, ms, 32)
        return {names[i].decode(): float(ms[i]) for i in range(n)}

    def debug_read(self, name, offset, nbytes, dtype=np.uint8):
        buf = np.zeros(nbytes, np.uint8)
        k = lib().phy_debug_read(self._h, name.encode(), offset, buf.ctypes.data, nbytes)
        if k < 0:
            raise self._err(int(k))
        return buf[:k].view(dtype)


def payloads(descs, out):
    return [out[d.out_off:d.out_off + d.out_len].tobytes() for d in descs]


def make_block_header(wrid, bewr, bhs, beso, bcss, sbol):
    s = np.ascontiguousarray(sbol, np.uint32)
    buf = np.zeros(4096, np.uint8)
    k = lib().phy_make_block_header(wrid, bewr, bhs, beso, bcss, s.ctypes.data, s.size, buf.ctypes.data, buf.size)
    if k == 0:
        raise PhyError(-7, "header does not fit")
    return buf[:k].tobytes()


def make_footer(np_ranks, fastq_size, n_blocks, n_subblocks, overlaps, block_order, lb_sizes):
    ov = np.ascontiguousarray(overlaps, np.int32); bo = np.ascontiguousarray(block_order, np.int32)
    lb = np.ascontiguousarray(lb_sizes, np.uint32)
    buf = np.zeros(64 + 4 * (bo.size + 2 * np_ranks), np.uint8)
    k = lib().phy_make_footer(np_ranks, fastq_size, n_blocks, n_subblocks, ov.ctypes.data, bo.ctypes.data, lb.ctypes.data,
                              buf.ctypes.data, buf.size)
    if k < 0:
        raise PhyError(k, lib().phy_strerror(k).decode())
    return buf[:k].tobytes()


def decode_subblock(payload, cap=None):
    """FASTQ text (uint8 array) of one subblock payload, decoded by phy_decode_subblock (host code, no device needed)."""
    p = np.ascontiguousarray(np.frombuffer(payload, np.uint8) if not isinstance(payload, np.ndarray) else payload)
    cap = int(cap) if cap else 16 * p.size + 4096
    while True:
        out = np.empty(cap, np.uint8)
        n = lib().phy_decode_subblock(p.ctypes.data, p.size, out.ctypes.data, out.size)
        if n == -5 and cap < (1 << 31):  # PHY_ERR_CAPACITY
            cap *= 4
            continue
        if n < 0:
            raise PhyError(int(n), "phy_decode_subblock")
        return out[:n]
