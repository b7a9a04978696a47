"""The oracle against the reference itself, compiled unmodified into oracle/_ref (oracle/Makefile).
Only meaningful where /root/reference exists (the build container); skipped elsewhere."""
import os

import numpy as np
import pytest

from oracle import phy_oracle as O
from phyngsc_b200 import container, synth

pytestmark = pytest.mark.skipif(not (O.have_reference() and os.path.exists(O.REF_KAT)), reason="oracle/_ref not built")


def _rand_freqs(rng, n):
    kind = rng.integers(0, 6)
    if kind == 0:
        f = rng.integers(0, 5, n)
    elif kind == 1:
        f = rng.integers(0, 2, n) * rng.integers(1, 1000, n)
    elif kind == 2:
        f = np.zeros(n, np.int64); f[rng.integers(0, n)] = rng.integers(1, 100)
    elif kind == 3:
        f = np.full(n, rng.integers(1, 9))
    elif kind == 4:
        f = (2 ** rng.integers(0, 20, n)).astype(np.int64)
    else:
        f = rng.integers(0, 100000, n)
    return f.astype(np.uint32)


def test_huffman_known_answers_against_reference_classes(oracle):
    rng = np.random.default_rng(5)
    sizes = [1, 2, 3, 4, 5, 7, 8, 16, 31, 41, 64, 100, 255, 256, 300, 512]
    for it in range(400):
        n = sizes[it % len(sizes)]
        f = _rand_freqs(rng, n)
        if n > 2 and not f.any():
            f[0] = 1
        c0, l0, t0 = O.ref_huffman(f, True)
        if l0.max() > 32:
            continue
        c1, l1, t1 = oracle.huffman(f, True)
        assert (c0 == c1).all() and (l0 == l1).all() and t0 == t1, (n, f.tolist())


@pytest.mark.parametrize("shape,mb,npr", [("36bp", 9, 2), ("100bp", 20, 2), ("100bp_huffdna", 6, 3), ("150bp_paired", 6, 4),
                                          ("var50_205", 12, 2), ("title_stress", 5, 2), ("degrade", 2, 2), ("mixed_amb", 3, 5),
                                          ("100bp", 60, 2)])
def test_rank_payloads_and_blocks_match_reference(shape, mb, npr, tmp_path, oracle):
    data = synth.fastq(shape, 100 + mb, target_bytes=mb * 1_000_000 + 4321)
    src = tmp_path / "in.fastq"
    data.tofile(src)
    oracle.run_reference(str(src), str(tmp_path / "out.ngsc"), np_ranks=npr, threads=1)
    ng = container.read_ngsc(str(tmp_path / "out.ngsc"))
    for r in range(npr):
        mine = oracle.compress_rank(data, npr, r)
        assert ng["per_rank_subblocks"][r] == mine["subblocks"]
        assert [b["raw"] for b in ng["per_rank_blocks"][r]] == mine["blocks"]


@pytest.mark.parametrize("threads", [2, 4])
def test_thread_count_rule_matches_reference(threads, tmp_path, oracle):
    """no_threads > 1: only the reference's last thread applies the stop rule (phyNGSC.cpp:261-266, 303, 315), so the
    short final window of rank 0 keeps more records than with one thread on this input.  The multi-threaded reference
    is flaky: it gets a few attempts with a short timeout."""
    data = synth.fastq("36bp", 702, target_bytes=40_000_000 + 977)
    src = tmp_path / "in.fastq"
    data.tofile(src)
    ng = None
    for _ in range(4):
        try:
            oracle.run_reference(str(src), str(tmp_path / "out.ngsc"), np_ranks=2, threads=threads, timeout=60)
            ng = container.read_ngsc(str(tmp_path / "out.ngsc"))
            break
        except Exception:  # noqa: BLE001
            continue
    if ng is None:
        pytest.skip("the reference did not finish with this thread count")
    mine = [oracle.compress_rank(data, 2, r, threads=threads) for r in range(2)]
    for r in range(2):
        assert ng["per_rank_subblocks"][r] == mine[r]["subblocks"]
        assert [b["raw"] for b in ng["per_rank_blocks"][r]] == mine[r]["blocks"]
    assert mine[0]["subblocks"] != oracle.compress_rank(data, 2, 0, threads=1)["subblocks"]


@pytest.mark.parametrize("shape", ["36bp", "100bp", "100bp_huffdna", "mixed_amb", "title_stress"])
def test_oracle_payloads_decode_with_the_reference_decoder(shape):
    """Second, independent pin: the oracle's payloads, fed to the reference's own Fetch* functions
    (tasks.cpp:625-1101 through oracle/ref_kat.cpp), give back the input FASTQ byte for byte.  (The reference's
    decoder cannot read what its own encoder writes for some title shapes -- SURVEY Q3 -- so those are not listed.)"""
    if not hasattr(O.ref_kat(), "ref_decode_subblock"):
        pytest.skip("oracle/_ref/libphyref_kat.so predates the decode hook")
    data = synth.fastq(shape, 77, target_bytes=1_200_000)
    r = O.compress_rank(data, 1, 0, window_bytes=300 * 1024)
    pos = 0
    for sb in r["subblocks"]:
        dec = O.ref_decode_subblock(sb, 16 * len(sb) + 4096)
        assert np.array_equal(dec, data[pos:pos + dec.size])
        pos += dec.size
    assert pos == data.size
