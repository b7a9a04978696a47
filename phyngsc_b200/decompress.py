"""phyNGSD -- decompressor for .ngsc files (SURVEY section 8(f) "next #2"; the reference names a phyNGSD.cpp in its
Makefile but does not ship it).

    python -m phyngsc_b200.decompress in.ngsc out.fastq [threads]

The container is read with phyngsc_b200.container (block headers, split subblocks, footer: tasks.cpp:1104-1200); every
subblock payload is decoded by phy_decode_subblock (host/phy_decode.hpp behind the C ABI) on a thread pool -- subblocks
are independent -- and written in rank order, which is file order of the original FASTQ.
"""
import sys
from concurrent.futures import ThreadPoolExecutor

from . import api, container


def decompress(ngsc_path, fastq_path, threads=8):
    ng = container.read_ngsc(ngsc_path)
    payloads = [sb for rank in ng["per_rank_subblocks"] for sb in rank]
    with ThreadPoolExecutor(max_workers=max(1, threads)) as ex, open(fastq_path, "wb") as f:
        total = 0
        for text in ex.map(api.decode_subblock, payloads):  # ctypes releases the GIL; results arrive in order
            f.write(text.tobytes())
            total += text.size
    return total, len(payloads)


def main(argv):
    if len(argv) < 3:
        print(__doc__)
        return 2
    n, k = decompress(argv[1], argv[2], int(argv[3]) if len(argv) > 3 else 8)
    print(f"{argv[1]}: {k} subblocks -> {n} bytes of FASTQ")
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv))
