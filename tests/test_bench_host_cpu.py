"""Host logic of bench.py that needs no GPU: the repeated-segment file image every rank cuts its working region from, the
record statistics behind the per-kernel algorithmic bytes, and the N-rank partition of one file (phyNGSC.cpp:113-124) -- the
oracle compressing the regions rank by rank must give the same subblocks as it gives on the whole image."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from phyngsc_b200 import api  # noqa: E402


def test_file_image_fill_is_the_tiled_segment(monkeypatch):
    monkeypatch.setattr(bench, "SEGMENT_MB", 1)
    img = bench.FileImage("100bp", 3)  # 3 MB image = a 1 MB segment three times
    seg = img.segment
    assert img.tiles == 3 and img.size == 3 * seg.size and seg[-1] == 10
    whole = np.tile(seg, 3)
    for start, end in ((0, 10), (seg.size - 5, seg.size + 7), (123_457, 2 * seg.size + 999), (0, img.size)):
        out = np.empty(end - start, np.uint8)
        assert np.array_equal(img.fill(start, end, out), whole[start:end])


def test_record_stats_counts_lines():
    data = np.frombuffer(b"@r1 x\nACGT\n+\nIIII\n@r2 yy\nAC\n+\nII\n", np.uint8)
    nrec, title, seq = bench.record_stats(data)
    assert (nrec, title, seq) == (2, 6 + 7, 4 + 2)


def test_ranks_of_one_image_tile_the_file(oracle, monkeypatch):
    """The strong-scaling partition bench.py measures: rank r gets region_slice(size, N, r) of the one image and compresses
    it with region_params(size, N, r); decoded back, the ranks' subblocks are the file, every record exactly once."""
    monkeypatch.setattr(bench, "SEGMENT_MB", 1)
    img = bench.FileImage("36bp", 2)
    whole = np.tile(img.segment, img.tiles)
    for n in (1, 2, 3):
        text = []
        for r in range(n):
            start, end = api.region_slice(img.size, n, r, slack=4096)
            region = img.fill(start, end, np.empty(end - start, np.uint8))
            assert np.array_equal(region, whole[start:end])
            for sb in oracle.compress_rank(whole, n, r, window_bytes=256 * 1024)["subblocks"]:
                text.append(api.decode_subblock(sb).tobytes())
        assert b"".join(text) == whole.tobytes()  # every record of the file exactly once, in order
