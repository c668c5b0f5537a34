import numpy as np
import pytest


def get_gsi():
    import gsi_b200
    return gsi_b200


@pytest.fixture(scope="session")
def gsi():
    import torch  # noqa: F401  (only to make sure CUDA is initialised the same way bench does)
    g = get_gsi()
    g.default_context()      # raises loudly (NoDeviceError) if the CUDA path is unavailable
    return g


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300))
