import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)

    return load


@pytest.fixture(scope="session")
def config1_corpus():
    """BASELINE configs[0] audio (10 speakers x 30 utterances x 3 s) regenerated from the seeds and split the way
    GMM_UBM.py:125 does; tests/golden/config1.npz holds the CRC of the test half."""
    from sklearn.model_selection import train_test_split

    from speech_signal_processing_b200 import synth

    x, y = synth.synth_corpus(10, 30, 48000)
    return train_test_split(x, y, test_size=0.3, random_state=0)
