import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def toy():
    """The reference's 3-tank test problem + its golden vectors (tests/golden/make_golden.py)."""
    from rapidnet_b200.problem import problem_from_npz_dict
    z = dict(np.load(os.path.join(ROOT, "tests", "golden", "toy.npz"), allow_pickle=False))
    prob = problem_from_npz_dict(z)
    engine = {k[len("engine."):]: v for k, v in z.items() if k.startswith("engine.")}
    smpc = {k[len("smpc."):]: v for k, v in z.items() if k.startswith("smpc.")}
    return prob, engine, smpc
