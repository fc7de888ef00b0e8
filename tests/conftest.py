import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def seed():
    """World name "test" (the reference's own test seed, generator.rs:157) -> u64."""
    import oracle

    return oracle.hash_world_name("test")


@pytest.fixture(scope="session")
def ctx(seed):
    """A vx context on cuda:0 (gpu tests only). Fails loudly if the extension is missing."""
    import voxel_game_b200 as vg

    c = vg.Context(0, seed=seed)
    yield c
    c.close()
