import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# Constants.py:7-20 as the positional arguments of VFE_preprocessing (Predict.py:21-28)
REF_ARGS = dict(xSize=0.5, ySize=0.25, zSize=0.25, sampleSize=35, maxVoxelX=100, maxVoxelY=200, maxVoxelZ=8)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with `pytest -m gpu` under gpurun")


@pytest.fixture(scope="session")
def ref_args():
    return dict(REF_ARGS)


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN
