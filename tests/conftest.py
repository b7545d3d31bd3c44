import glob
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "lidar-gs_b200"), os.path.join(ROOT, "oracle"), ROOT, os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


GOLDENS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "g[0-9]*.npz")))          # 3-D rasterizer
SURFEL_GOLDENS = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "gs[0-9]*.npz")))  # surfel rasterizer


@pytest.fixture(params=GOLDENS, ids=[os.path.basename(g)[:-4] for g in GOLDENS])
def golden(request):
    import util
    return util.load_golden(request.param)


@pytest.fixture(params=SURFEL_GOLDENS, ids=[os.path.basename(g)[:-4] for g in SURFEL_GOLDENS])
def surfel_golden(request):
    import util
    return util.load_golden(request.param)
