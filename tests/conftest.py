import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def mm():
    """The product package; the in-tree library is built on demand (CPU box: nvcc cross-compiles sm_100a)."""
    import mirror_maze_b200 as pkg

    if not os.path.exists(pkg.library_path()):
        import __graft_entry__

        __graft_entry__.build()
    pkg.load_library()
    return pkg


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as o

    o.lib()
    return o


@pytest.fixture(scope="session")
def noise(mm):
    return mm.load_noise()


@pytest.fixture(scope="session")
def scenes(mm):
    cache = {}

    def get(n, seed=0):
        if (n, seed) not in cache:
            cache[(n, seed)] = mm.MazeScene(n, seed)
        return cache[(n, seed)]

    return get


@pytest.fixture(scope="session")
def renderer(mm):
    r = mm.Renderer(0)
    yield r
    r.close()
