import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def small_scene():
    from model_matching_b200 import synth
    sc = synth.make_scene(n_points=40000, extent=(0.7, 0.5, 0.5), n_objects=5, seed=7)
    mpos, mnrm = synth.make_model(256)
    return sc, mpos, mnrm


@pytest.fixture(scope="session")
def gpu_ctx():
    from model_matching_b200 import Context
    ctx = Context(0)
    yield ctx
    ctx.close()
