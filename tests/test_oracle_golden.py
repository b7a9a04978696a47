"""The oracle (oracle/phy_oracle.c) against golden .ngsc files minted from the unmodified reference
(tests/golden/make_golden.py).  Runs anywhere (no /root/reference, no GPU)."""
import hashlib
import json
import os

import pytest

from phyngsc_b200 import container, synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MANIFEST = json.load(open(os.path.join(GOLD, "manifest.json")))


@pytest.mark.parametrize("case", MANIFEST, ids=[c["file"] for c in MANIFEST])
def test_oracle_reproduces_reference_file(case, oracle):
    data = synth.fastq(case["shape"], case["seed"], target_bytes=case["target_bytes"])
    assert hashlib.sha256(data.tobytes()).hexdigest() == case["input_sha256"], "synthetic generator drifted"
    ng = container.read_ngsc(os.path.join(GOLD, case["file"]))
    npr = case["np"]
    assert ng["footer"]["np"] == npr and ng["footer"]["fastq_size"] == data.size
    ranks = [oracle.compress_rank(data, npr, r) for r in range(npr)]
    for r in range(npr):
        assert ng["per_rank_subblocks"][r] == ranks[r]["subblocks"]          # every payload, bit-exact
        assert [b["raw"] for b in ng["per_rank_blocks"][r]] == ranks[r]["blocks"]  # headers + block split
    # footer for the block order the reference happened to write
    foot = oracle.make_footer(npr, data.size, sum(len(x["blocks"]) for x in ranks), sum(len(x["subblocks"]) for x in ranks),
                              [x["wr_overlap"] for x in ranks], ng["footer"]["block_order"],
                              [x["last_block_size"] for x in ranks])
    assert foot == ng["footer_bytes"]
