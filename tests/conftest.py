import os
import sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def make_streams(R, C, d, nsteps, seed, N=None, mult=40):
    """Replay streams for R ranks: normals Z, uniforms U in [0,1), ints I in [0,N)."""
    N = R * C if N is None else N
    rng = np.random.default_rng(seed)
    Z = rng.standard_normal((R, nsteps * C * d * mult))
    U = rng.random((R, nsteps * C * mult))
    I = rng.integers(0, N, (R, nsteps * C * mult), dtype=np.int32)
    return Z, U, I


def tiled_pinit(C, d):
    """The reference mains' four starting points (mcpar-rosen1.cc:43), tiled over chains
    and repeated d/2 times across parameters."""
    base = np.array([[0.0, 0.0], [2.0, 2.0], [0.0, 1.5], [0.0, -2.0]])
    p = np.tile(base, ((C + 3) // 4, max(1, d // 2)))[:C, :d]
    return np.ascontiguousarray(p)


@pytest.fixture(scope="session")
def mcgpu_lib():
    from mcpar_b200 import engine
    return engine.load()
