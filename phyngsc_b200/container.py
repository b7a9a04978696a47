"""Reader for the .ngsc container (block headers, split subblocks, footer).

Mirrors the layout written by the reference's MakeHeader / MakeFooter (tasks.cpp:1104-1200,
structures.h:310-333, phyNGSC.cpp:842-928).  Used by the block-keyed comparator: the reference's
block order in the file is non-deterministic (shared file pointer), so files are compared per
(WRID, ordinal-within-WRID), never byte-for-byte as a whole.
"""
import math


class BitReader:
    def __init__(self, data, pos=0):
        self.d = data
        self.bit = pos * 8

    def get(self, n):
        v = 0
        for _ in range(n):
            byte = self.d[self.bit >> 3]
            v = (v << 1) | ((byte >> (7 - (self.bit & 7))) & 1)
            self.bit += 1
        return v

    def align(self):
        self.bit = (self.bit + 7) & ~7

    @property
    def byte_pos(self):
        return self.bit >> 3


def ceil_log2(x):
    return 0 if x <= 1 else (x - 1).bit_length()


def parse_footer(data):
    """-> dict with np, fastq_size, n_blocks, n_subblocks, overlaps[1..], block_order, lb_sizes, footer_len."""
    flen = (data[-2] << 8) | data[-1]
    start = len(data) - 2 - flen
    r = BitReader(data, start)
    BEPS, BEFS, BEBS, BESS, BELB, BEOV, LBES = r.get(4), r.get(6), r.get(4), r.get(4), r.get(5), r.get(4), r.get(1)
    np_ = r.get(BEPS)
    fs = (r.get(BEFS - 32) << 32) | r.get(32) if BEFS > 32 else r.get(BEFS)
    nb, ns = r.get(BEBS), r.get(BESS)
    ov = [0] + [r.get(BEOV) for _ in range(np_ - 1)]
    cb = ceil_log2(np_)
    order = [r.get(cb) for _ in range(nb)]
    lbs = [r.get(BELB) for _ in range(np_)] if not LBES else None
    return dict(np=np_, fastq_size=fs, n_blocks=nb, n_subblocks=ns, overlaps=ov, block_order=order, lb_sizes=lbs,
                lbes=LBES, footer_start=start, footer_len=flen + 2)


def parse_block_header(data, pos, np_ranks):
    r = BitReader(data, pos)
    wrid = r.get(ceil_log2(np_ranks))
    bhs, nosb, beso, bcss = r.get(12), r.get(6), r.get(5), r.get(2)
    sbol = [r.get(beso) for _ in range(nosb)]
    r.align()
    return dict(wrid=wrid, bhs=bhs, nosb=nosb, beso=beso, bcss=bcss, sbol=sbol, header_len=r.byte_pos - pos)


def parse_blocks(data, np_ranks, block_bytes=1 << 23, end=None):
    """Walk the blocks of a .ngsc image in file order.  A block is block_bytes long unless it is the
    last block of its rank, in which case it is header + sum(SBOL).  Returns a list of dicts with
    offset/length/header fields/raw bytes."""
    end = len(data) if end is None else end
    pos, out = 0, []
    while pos < end:
        h = parse_block_header(data, pos, np_ranks)
        body = sum(h["sbol"])
        length = h["bhs"] + body
        if length > block_bytes or h["bhs"] != h["header_len"]:
            raise ValueError(f"bad block header at {pos}: {h}")
        h.update(offset=pos, length=length, raw=bytes(data[pos:pos + length]))
        out.append(h)
        pos += length
    if pos != end:
        raise ValueError("blocks do not tile the file")
    return out


def blocks_by_rank(blocks, np_ranks):
    per = [[] for _ in range(np_ranks)]
    for b in blocks:
        per[b["wrid"]].append(b)
    return per


def subblocks_of_rank(rank_blocks):
    """Re-join subblocks split across consecutive blocks of one rank (LSBS = bit0, FSBS = bit1)."""
    subs, carry = [], None
    for b in rank_blocks:
        p = b["bhs"]
        raw = b["raw"]
        for i, n in enumerate(b["sbol"]):
            piece = raw[p:p + n]; p += n
            first, last = i == 0, i == len(b["sbol"]) - 1
            if first and carry is not None:
                piece = carry + piece
                carry = None
            if last and (b["bcss"] & 1):
                carry = piece
            else:
                subs.append(piece)
    if carry is not None:
        raise ValueError("dangling split subblock")
    return subs


def read_ngsc(path_or_bytes, np_ranks=None, block_bytes=1 << 23):
    data = path_or_bytes if isinstance(path_or_bytes, (bytes, bytearray, memoryview)) else open(path_or_bytes, "rb").read()
    foot = parse_footer(data)
    npr = foot["np"] if np_ranks is None else np_ranks
    blocks = parse_blocks(data, npr, block_bytes, end=foot["footer_start"])
    per = blocks_by_rank(blocks, npr)
    return dict(footer=foot, blocks=blocks, per_rank_blocks=per, per_rank_subblocks=[subblocks_of_rank(b) for b in per],
                footer_bytes=bytes(data[foot["footer_start"]:]))
