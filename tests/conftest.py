import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (B200); run with -m gpu')


@pytest.fixture(scope='session')
def lib():
    """The C-ABI library; built on demand (nvcc cross-compiles without a GPU)."""
    from nekstab_next_b200 import build, _capi
    build.build()
    return _capi.load()


@pytest.fixture(scope='session')
def ctx(lib):
    """A device context; GPU tests fail (not skip) when the CUDA path is unavailable."""
    import nekstab_next_b200 as nb
    c = nb.Context(device=0)
    yield c
    c.close()
