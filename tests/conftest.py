import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box with -m gpu)')


@pytest.fixture(scope='session')
def lib():
    """The C-ABI library, built in-tree if needed (nvcc cross-compiles without a GPU)."""
    from elisa_b200 import _lib, build

    build.build()
    return _lib.load()


@pytest.fixture(autouse=True)
def _chunk_pipeline_unless_asked(request, monkeypatch):
    """Small test problems would all take the one-kernel small-batch path (tiny_eval: <= 1024 rows on a small
    spectrum) and leave the chunk pipeline -- unfold, tensor-core folds, statistic, contraction: the path the
    bench measures -- untested at test sizes.  Every module but the one that tests the small-batch kernel
    itself therefore runs with it switched off (ELISA_B200_TINY is read at plan creation)."""
    if getattr(request.module, 'SMALL_BATCH_KERNEL', False):
        monkeypatch.delenv('ELISA_B200_TINY', raising=False)
    else:
        monkeypatch.setenv('ELISA_B200_TINY', '0')
