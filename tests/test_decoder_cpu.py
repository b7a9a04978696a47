"""The subblock decoder (phyngsc_b200/host/phy_decode.hpp behind phy_decode_subblock) on the CPU: payloads written by
the oracle and the .ngsc files minted from the compiled reference decode back to the input FASTQ -- including the
title shapes whose streams the reference's own Fetch* functions cannot read."""
import hashlib
import json
import os

import numpy as np
import pytest

from phyngsc_b200 import api, container, decompress, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SHAPES = ["36bp", "100bp", "100bp_huffdna", "150bp_paired", "var50_205", "var50_250", "title_stress", "degrade", "mixed_amb"]


@pytest.mark.parametrize("shape", SHAPES)
def test_oracle_payloads_round_trip(shape, oracle):
    data = synth.fastq(shape, 91, target_bytes=1_500_000)
    r = oracle.compress_rank(data, 1, 0, window_bytes=384 * 1024)
    pos = 0
    for sb in r["subblocks"]:
        dec = api.decode_subblock(sb)
        assert np.array_equal(dec, data[pos:pos + dec.size])
        pos += dec.size
    assert pos == data.size


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLD, "manifest.json"))), ids=lambda c: c["file"])
def test_reference_ngsc_files_decompress_to_their_input(case, tmp_path):
    """Whole container path: footer, block headers, subblocks split across blocks, several ranks."""
    out = tmp_path / "out.fastq"
    n, k = decompress.decompress(os.path.join(GOLD, case["file"]), str(out), threads=2)
    got = out.read_bytes()
    assert n == case["input_bytes"] == len(got) and k >= case["np"]
    assert hashlib.sha256(got).hexdigest() == case["input_sha256"]


@pytest.mark.parametrize("case", json.load(open(os.path.join(GOLD, "manifest.json")))[::3], ids=lambda c: c["file"])
def test_cpp_decompressor_binary(case, tmp_path):
    """host/phyNGSD_b200 (the C++ program a user runs) on reference-minted files."""
    import subprocess
    from phyngsc_b200 import build
    exe = build.build_decompressor()
    out = tmp_path / "out.fastq"
    p = subprocess.run([exe, os.path.join(GOLD, case["file"]), str(out), "3"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert hashlib.sha256(out.read_bytes()).hexdigest() == case["input_sha256"]
    bad = tmp_path / "bad.ngsc"
    bad.write_bytes(open(os.path.join(GOLD, case["file"]), "rb").read()[:-7])
    assert subprocess.run([exe, str(bad), str(out)], capture_output=True, timeout=120).returncode != 0


def test_garbage_is_an_error_not_a_crash():
    rng = np.random.default_rng(5)
    for n in (0, 3, 19, 200, 5000):
        with pytest.raises(api.PhyError):
            api.decode_subblock(rng.integers(0, 256, n, dtype=np.uint8))
    good = synth.fastq("36bp", 3, target_bytes=100_000)
    assert container is not None and good.size > 0


def test_cpp_decompressor_file_without_trailing_newline(tmp_path, oracle):
    """A FASTQ whose last line has no newline: the encoder sees a virtual one, the decoder writes one, the footer knows the
    real size -- the decompressor drops the extra byte instead of reporting a size mismatch.  The container is put together
    from the oracle's blocks and footer (rank-major order, as the driver writes it)."""
    import subprocess
    from phyngsc_b200 import build
    data = synth.fastq("100bp", 17, target_bytes=700_000)
    assert data[-1] == 10
    data = data[:-1]
    ranks = [oracle.compress_rank(data, 2, r, window_bytes=256 * 1024) for r in range(2)]
    order = [r for r in range(2) for _ in ranks[r]["blocks"]]
    foot = oracle.make_footer(2, data.size, len(order), sum(len(x["subblocks"]) for x in ranks), [x["wr_overlap"] for x in ranks], order,
                              [x["last_block_size"] for x in ranks])
    src, out = tmp_path / "in.ngsc", tmp_path / "out.fastq"
    src.write_bytes(b"".join(b for r in range(2) for b in ranks[r]["blocks"]) + foot)
    p = subprocess.run([build.build_decompressor(), str(src), str(out), "2"], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0, p.stdout + p.stderr
    assert out.read_bytes() == data.tobytes()


def test_damaged_lengths_do_not_size_buffers():
    """A constant title token whose length word is damaged must fail on the truncation check, not allocate gigabytes."""
    data = synth.fastq("36bp", 3, target_bytes=60_000)
    from oracle import phy_oracle
    sb = bytearray(phy_oracle.compress_rank(data, 1, 0)["subblocks"][0])
    # info: R, max_qlen, max_slen words, 3 bytes, flags word, length bits; the title header follows: nf word, then per field
    # sep byte, constant byte, (constant:) length word.  Field 0 ("@ERR...") is constant: blow up its length word.
    R = int.from_bytes(sb[0:4], "big"); mq = int.from_bytes(sb[4:8], "big")
    off = 19 + (R * mq.bit_length() + 7) // 8 + 4
    assert sb[off + 1] == 1  # field 0 is constant
    sb[off + 2:off + 6] = (0xF0000000).to_bytes(4, "big")
    with pytest.raises(api.PhyError):
        api.decode_subblock(np.frombuffer(bytes(sb), dtype=np.uint8))
