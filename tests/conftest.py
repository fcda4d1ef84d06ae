import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "cpecan-signal_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def zymo():
    from cpecan_signal import synth
    g = dict(np.load(os.path.join(GOLDEN, "zymo_golden.npz")))
    g["ref"] = open(os.path.join(GOLDEN, "ZymoRef.txt")).readline().strip()
    g["read"] = synth.load_npread(os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"))
    return g


@pytest.fixture(scope="session")
def syn_golden():
    return dict(np.load(os.path.join(GOLDEN, "synthetic_golden.npz")))


@pytest.fixture(scope="session")
def template_tables():
    from cpecan_signal import synth
    return synth.load_model_file(synth.TEMPLATE_MODEL)


@pytest.fixture(scope="session")
def engine():
    from cpecan_signal import Engine
    eng = Engine(0)
    yield eng
    eng.close()
