import os
import sys

import pytest

# The multi-rank tests run every rank of a process grid as a host thread on ONE GPU, each with its own streams, and the
# inter-rank barrier is a spinning kernel.  Streams are multiplexed onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (8 by
# default): with more streams than queues a spinning barrier kernel can sit in front of the very kernel (another rank's
# arrival) it waits for.  One process per GPU -- the product's deployment -- never gets there; the single-GPU harness needs
# enough queues for 16 ranks x (1 compute + copy lanes).  Must be set before the CUDA context exists.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def cuda_device():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    torch.cuda.set_device(0)
    return torch.device("cuda:0")
