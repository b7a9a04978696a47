import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


# Shapes of the parity regression set (SURVEY.md Appendix A preamble + section 8(d)).
SHAPES = ["36bp", "100bp", "100bp_huffdna", "150bp_paired", "var50_205", "title_stress", "degrade", "mixed_amb"]


@pytest.fixture(scope="session")
def oracle():
    from oracle import phy_oracle
    phy_oracle.build()
    return phy_oracle
