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
