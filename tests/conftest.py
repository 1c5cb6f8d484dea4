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
    """The unmodified reference kernels (oracle/_ref).  On a GPU box their absence is an ERROR, not a skip: every
    reference comparison of the -m gpu suite would silently vanish (oracle/_ref is built by oracle/Makefile here, where
    /root/reference exists, and travels with the snapshot).  Without a GPU (no -m gpu test runs) it is None."""
    import ref_gpu
    import torch
    if ref_gpu.available():
        return ref_gpu
    if torch.cuda.is_available():
        pytest.fail("oracle/_ref/libsph3d_ref.so is missing on a GPU box: run `make -C oracle ref` where /root/reference "
                    "exists (python -c 'import __graft_entry__ as g; g.build()') before shipping the snapshot")
    return None


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
