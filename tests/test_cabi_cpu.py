"""CPU-side checks of the C ABI library: it loads, exports every symbol include/phyngsc_b200.h declares,
its host-only helpers (block header / footer / first-record sync) agree with the oracle, and compute entry
points fail loudly without a GPU instead of falling back."""
import os
import re

import numpy as np
import pytest

from phyngsc_b200 import api, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "phyngsc_b200.h")).read()
    declared = set(re.findall(r"\b(phy_[a-z_0-9]+)\s*\(", hdr))
    assert declared == set(api.EXPORTS)
    L = api.lib()
    for name in declared:
        assert hasattr(L, name), name
    assert L.phy_abi_version() == 3


def test_block_header_and_footer_match_oracle(oracle):
    rng = np.random.default_rng(3)
    for _ in range(200):
        npr = int(rng.integers(2, 40))
        bewr = (npr - 1).bit_length()
        nosb = int(rng.integers(1, 60))
        sbol = rng.integers(1, 1 << 22, nosb).astype(np.uint32)
        beso = int(sbol.max()).bit_length() + int(rng.integers(0, 2))
        bhs = (bewr + 18 + beso * nosb + 7 + 7) // 8
        wrid, bcss = int(rng.integers(0, npr)), int(rng.integers(0, 4))
        buf = np.zeros(4096, np.uint8)
        k = oracle.lib().phy_oracle_make_header(wrid, bewr, bhs, beso, bcss, sbol.ctypes.data, nosb, buf.ctypes.data, buf.size)
        assert api.make_block_header(wrid, bewr, bhs, beso, bcss, sbol) == buf[:k].tobytes()
        nblk = int(rng.integers(1, 300))
        ov = [0] + [int(x) for x in rng.integers(0, 400, npr - 1)]
        if max(ov) == 0:
            ov[-1] = 7
        order = rng.integers(0, npr, nblk)
        lbs = rng.integers(1, 1 << 23, npr) if rng.integers(0, 2) else np.full(npr, 12345)
        fs = int(rng.integers(1, 1 << 40))
        args = (npr, fs, nblk, int(rng.integers(1, 30000)), ov, order, lbs)
        assert api.make_footer(*args) == oracle.make_footer(*args)


def test_first_record_sync_matches_oracle(oracle):
    data = synth.fastq("100bp", 9, target_bytes=400_000)
    for npr in (2, 3, 5, 7):
        for r in range(1, npr):
            start = r * (data.size // npr)
            reg = data[start:]
            got = api.lib().phy_find_first_record(reg.ctypes.data, reg.size)
            assert got == oracle.compress_rank(data, npr, r)["wr_overlap"]


def test_no_cpu_fallback_without_a_gpu():
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(api.PhyError) as e:
        api.Context(0)
    assert e.value.code == -6
