import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


def _built(path):
    return os.path.exists(os.path.join(ROOT, path))


@pytest.fixture(scope="session", autouse=True)
def _build_cpu_libs():
    """CPU-side libraries are cheap to build; the CUDA library must already exist for -m gpu."""
    import subprocess
    if not (_built("oracle/liboracle.so") and _built("minimaloptix_b200/libmox_host.so")):
        subprocess.check_call(["make", "-C", ROOT, "oracle", "host"], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def orc():
    import oracle
    return oracle


@pytest.fixture(scope="session")
def host():
    from minimaloptix_b200 import host as h
    return h


@pytest.fixture(scope="session")
def gpu_backend():
    import minimaloptix_b200 as mox
    return mox.gpu()  # raises if libmox.so is missing: no CPU fallback


@pytest.fixture(scope="session")
def api_tables(host, orc):
    import minimaloptix_b200 as mox

    class T:
        oracle = host.ApiTable(orc.ORACLE_LIB, "orc_")
        _gpu = None

        @property
        def gpu(self):
            if self._gpu is None:
                self._gpu = host.ApiTable(mox.GPU_LIB, "mox_")
            return self._gpu
    return T()


def random_rays(n, lo, hi, seed, tmin=1e-3, tmax=1e27):
    """Incoherent rays: origins uniform in the box [lo, hi], directions uniform on the sphere."""
    rng = np.random.default_rng(seed)
    o = rng.uniform(lo, hi, size=(n, 3)).astype(np.float32)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d = d.astype(np.float32)
    # renormalise in float32 the way the render path does (v * (1/sqrt(dot)))
    inv = (np.float32(1.0) / np.sqrt((d * d).sum(axis=1, dtype=np.float32))).astype(np.float32)
    d = (d * inv[:, None]).astype(np.float32)
    rays = np.empty((n, 8), dtype=np.float32)
    rays[:, 0:3] = o
    rays[:, 3] = tmin
    rays[:, 4:7] = d
    rays[:, 7] = tmax
    return rays
