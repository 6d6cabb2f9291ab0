import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gaussian-pcloud-render_b200"), os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle32():
    from oracle.oracle import Oracle
    return Oracle(32)


@pytest.fixture(scope="session")
def oracle64():
    from oracle.oracle import Oracle
    return Oracle(64)


@pytest.fixture(scope="session")
def golden():
    """name -> (case dict, rasterizer kwargs, npz of reference outputs)."""
    from make_golden import golden_cases, input_checksum, rast_kwargs
    out = {}
    for name, case in golden_cases().items():
        kw = rast_kwargs(case)
        g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
        assert str(g["checksum"]) == input_checksum(kw), f"golden inputs of {name} drifted"
        out[name] = (case, kw, g)
    return out


GOLDEN_NAMES = ["tiny_sh3", "tiny_ties", "human_m13", "precomp", "cull_edges", "sh0_packed"]
