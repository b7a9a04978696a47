"""In-tree build of the native pieces (nvcc / g++ directly; no JIT cache).

  csrc/libphyngsc_b200.so   the product: sm_100a kernels + C ABI (include/phyngsc_b200.h)
  csrc/libphysynth.so       synthetic FASTQ generator (plain C)
  host/phyNGSC_b200         drop-in host driver (C++/MPI; built against the fork-based mpi.h stand-in
                            when no MPI is installed)
  host/phyNGSD_b200         decompressor (plain C++)
"""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
HOST = os.path.join(HERE, "host")
LIB = os.path.join(CSRC, "libphyngsc_b200.so")
DRIVER = os.path.join(HOST, "phyNGSC_b200")
DECOMP = os.path.join(HOST, "phyNGSD_b200")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--extended-lambda",
              "-Xcompiler", "-fPIC", "-shared"]


def _newer(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def nvcc():
    return shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"


def build_lib(force=False, verbose=False):
    src = [os.path.join(CSRC, f) for f in ("phy_b200.cu", "phy_container.cpp")]
    deps = src + [os.path.join(CSRC, f) for f in ("phy_core.cuh", "phy_kernels.cuh", "phy_fast.cuh", "phy_encode.cuh", "phy_seqstat.cuh", "phy_title.cuh")] + [os.path.join(HERE, "..", "include", "phyngsc_b200.h"), os.path.join(HOST, "phy_decode.hpp")]
    if force or _newer(LIB, deps):
        extra = os.environ.get("PHY_NVCC_EXTRA", "").split()  # experiment switches (-DPHY_...=n); empty in normal builds
        cmd = [nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB] + src
        subprocess.check_call(cmd)
    return LIB


def build_driver(force=False):
    src = os.path.join(HOST, "phyNGSC_b200.cpp")
    if not os.path.exists(src):
        return None
    if force or _newer(DRIVER, [src, LIB]):
        mpicxx = shutil.which("mpicxx")
        cxx = [mpicxx] if mpicxx else ["g++", "-I" + os.path.join(HOST, "mpi_shim")]
        subprocess.check_call(cxx + ["-O2", "-std=c++17", "-I" + os.path.join(HERE, "..", "include"), "-o", DRIVER, src,
                                     "-L" + CSRC, "-lphyngsc_b200", "-Wl,-rpath," + CSRC, "-lpthread"])
    return DRIVER


def build_decompressor(force=False):
    """host/phyNGSD_b200: plain C++ (no CUDA, no MPI)."""
    src = [os.path.join(HOST, "phyNGSD_b200.cpp"), os.path.join(HOST, "phy_decode.hpp")]
    if force or _newer(DECOMP, src):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-Wall", "-o", DECOMP, src[0], "-lpthread"])
    return DECOMP


def build_all(force=False):
    from . import synth
    synth.build(force)
    build_lib(force)
    build_driver(force)
    build_decompressor(force)


if __name__ == "__main__":
    build_all(force=True)
    print("built", LIB)
