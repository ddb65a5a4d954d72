import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


@pytest.fixture(scope="session")
def O():
    """The CPU oracle (oracle/liboracle.so, built on demand)."""
    from oracle import oracle as orc
    orc.build()
    orc.lib()
    return orc


@pytest.fixture(scope="session")
def sb():
    import shems_b200
    return shems_b200


@pytest.fixture(scope="session")
def P98(O):
    return O.params_for_charger(98)


@pytest.fixture(scope="session")
def train_series(sb):
    return sb.series.synth_charger98(4320, seed=98)


@pytest.fixture(scope="session")
def charger98_test_series():
    return np.load(os.path.join(GOLDEN, "charger98_test_series.npz"))["series"]


def random_states(rng, n, series, P):
    """Random but plausible env states tied to random rows of `series` (+ adversarial corners)."""
    nrows = series.shape[1]
    idx = rng.integers(1, nrows, size=n).astype(np.int32)  # 1-based, idx+1 <= nrows
    obs = np.zeros((9, n), np.float32)
    obs[0] = rng.uniform(0, P.b_soc_max, n)
    obs[1:] = series[:, idx - 1]
    connected = obs[2] >= 0
    obs[1] = np.where(connected, rng.uniform(0, 1, n), 1.0)
    k = n // 8
    obs[0, :k] = rng.choice([0.0, 5e-4, 1e-3, 1.0000001e-3, P.b_soc_max, P.b_soc_max * 0.95, 0.011, 3.3], k)  # battery corners
    obs[1, k:2 * k] = np.where(connected[k:2 * k], rng.choice([0.0, 0.5, 0.98999995, 0.99, 0.99999994, 1.0], k), 1.0)
    return obs, idx
