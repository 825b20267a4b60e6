import importlib.util
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
for p in (str(ROOT), str(ROOT / "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def load_orbfront():
    """Import the (hyphen-named) product package as module `orbfront_b200`."""
    if "orbfront_b200" in sys.modules:
        return sys.modules["orbfront_b200"]
    spec = importlib.util.spec_from_file_location("orbfront_b200", ROOT / "adaptive-rgbd-localization-mappig_b200" / "__init__.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["orbfront_b200"] = mod
    spec.loader.exec_module(mod)
    return mod


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def orc():
    from oracle import oracle as o
    o.build()
    return o


@pytest.fixture(scope="session")
def ob():
    return load_orbfront()


@pytest.fixture(scope="session")
def texture():
    import synth
    return synth.make_texture(0, 480, 640)
