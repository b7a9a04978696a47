"""Mint golden .ngsc fixtures from the UNMODIFIED reference (oracle/_ref/phyNGSC_ref, built by
oracle/Makefile from /root/reference) -- run in the build container, where /root/reference exists:

    python tests/golden/make_golden.py

For every regression shape a small synthetic FASTQ (phyngsc_b200.synth, fixed seed) is compressed
by the reference at np=2 and np=3, threads=1; the resulting .ngsc files are committed next to a
manifest holding the input's sha256 (inputs are regenerated from the seed, not stored).
The reference ships no golden vectors of its own, so these files are what
pins oracle/phy_oracle.c on machines without /root/reference (the GPU box).
"""
import hashlib
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle import phy_oracle as O  # noqa: E402
from phyngsc_b200 import synth  # noqa: E402

CASES = [  # (shape, seed, target_bytes)
    ("36bp", 11, 300_000), ("100bp", 12, 400_000), ("100bp_huffdna", 13, 300_000), ("150bp_paired", 14, 300_000),
    ("var50_205", 15, 300_000), ("title_stress", 16, 250_000), ("degrade", 17, 120_000), ("mixed_amb", 18, 150_000),
]


def main():
    O.build()
    assert O.have_reference(), "oracle/_ref/phyNGSC_ref missing: needs /root/reference"
    manifest = []
    for shape, seed, nbytes in CASES:
        data = synth.fastq(shape, seed, target_bytes=nbytes + 77)  # +77: keep rank boundaries off record starts (Q12)
        tmp = f"/tmp/golden_{shape}.fastq"
        data.tofile(tmp)
        for npr in (2, 3):
            name = f"{shape}_np{npr}.ngsc"
            O.run_reference(tmp, os.path.join(HERE, name), np_ranks=npr, threads=1)
            manifest.append(dict(shape=shape, seed=seed, target_bytes=nbytes + 77, np=npr, file=name,
                                 input_sha256=hashlib.sha256(data.tobytes()).hexdigest(), input_bytes=int(data.size)))
        os.remove(tmp)
    with open(os.path.join(HERE, "manifest.json"), "w") as f:
        json.dump(manifest, f, indent=1)
    print(f"wrote {len(manifest)} golden files")


if __name__ == "__main__":
    main()
