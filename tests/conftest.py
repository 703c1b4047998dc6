import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


# Every workspace the engine allocates during the tests starts as NaN bytes: a kernel that reads workspace it
# (or an earlier kernel of the same step) never wrote shows up as a NaN instead of passing on lucky zeros.
os.environ.setdefault("CBFSSM_B200_POISON_WS", "1")
