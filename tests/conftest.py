import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def flat_model():
    from phase_guided_terrain_traversal_b200.model import compile_model
    return compile_model("flat_terrain")


@pytest.fixture(scope="session")
def stairs_model():
    from phase_guided_terrain_traversal_b200.model import compile_model
    return compile_model("stairs")


@pytest.fixture(scope="session")
def train_cfg():
    from phase_guided_terrain_traversal_b200.go2.configs import default_config, training_overrides
    return training_overrides(default_config())
