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
def hdp_fixture(tmp_path_factory):
    """The reference's serialised NanoporeHDP fixture (tests/test_hdp/testTemplate.nhdp, committed gzip-compressed), its
    goldens from the unmodified reference (oracle/make_golden.py hdp), and the fixture read's template events descaled
    as vanillaAlign does for this machine (only the first third: impl/nanopore.c:34-38)."""
    import gzip
    import shutil
    from cpecan_signal import hdp, synth
    d = tmp_path_factory.mktemp("hdp")
    path = str(d / "testTemplate.nhdp")
    with gzip.open(os.path.join(GOLDEN, "hdp", "testTemplate.nhdp.gz"), "rb") as fi, open(path, "wb") as fo:
        shutil.copyfileobj(fi, fo)
    g = dict(np.load(os.path.join(GOLDEN, "hdp", "zymo_hdp_golden.npz")))
    rd = synth.load_npread(os.path.join(GOLDEN, "ZymoC_ch_1_file1.npRead"))
    ev = np.array(rd["template_events"], dtype=np.float64).reshape(-1, 3).copy()
    k = (len(ev) + 2) // 3
    ev[:k, 0] = (ev[:k, 0] - rd["template_params"][1]) / rd["template_params"][0]
    return dict(path=path, hdp=hdp.load_nhdp(path), golden=g, events=ev)


@pytest.fixture(scope="session")
def engine():
    from cpecan_signal import Engine
    eng = Engine(0)
    yield eng
    eng.close()
