import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the oracle and the synthetic-input generator once (g++ only, seconds).  The CUDA library is built by
    __graft_entry__.build(); tests that need it fail loudly if it is missing."""
    from oracle import pq_oracle
    from pepquery_b200 import synth
    pq_oracle.build()
    synth.build()


@pytest.fixture(scope="session")
def small_inputs():
    from pepquery_b200 import synth
    prot = synth.proteins(300)
    tgt = synth.targets(25)
    sp = synth.spectra(700, tgt, prot)
    return prot, tgt, sp


def have_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False
