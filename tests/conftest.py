import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.dirname(os.path.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    import oracle as O
    O.build()
    return O


@pytest.fixture(scope="session")
def pkg():
    import sph3d_gcn_b200 as S
    return S


@pytest.fixture(scope="session")
def ref():
    """The unmodified reference kernels (oracle/_ref), or None when the .so did not travel."""
    import ref_gpu
    return ref_gpu if ref_gpu.available() else None


@pytest.fixture()
def tune(pkg):
    """set SPH3D_* launch tunables for one test: the library reads the environment once at load, so every change is
    followed by sph3d_reload_tunables(); the previous environment is restored (and re-read) afterwards."""
    saved = {}

    def set_(**kw):
        for k, v in kw.items():
            saved.setdefault(k, os.environ.get(k))
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = str(v)
        pkg._lib.reload_tunables()

    yield set_
    for k, v in saved.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    pkg._lib.reload_tunables()
