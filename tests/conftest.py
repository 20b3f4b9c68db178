import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    torch.set_num_threads(max(1, min(16, os.cpu_count() or 1)))


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(name):
        z = np.load(os.path.join(GOLDEN, name + ".npz"))
        return {k: torch.from_numpy(z[k]) if z[k].shape != () else torch.tensor(z[k].item()) for k in z.files}
    return load


@pytest.fixture(scope="session")
def synth_sd16():
    """The golden weights: synthetic_state_dict(seed=0, M=16)."""
    from robir_b200 import synthetic
    return synthetic.synthetic_state_dict(0, num_lgt_sgs=16)


@pytest.fixture(scope="session")
def oracle_octrees(synth_sd16):
    """(primary, secondary) oracle octrees for the golden weights; built once per session (~15 s on 8 cores)."""
    import copy
    import robir_oracle as O
    import tracers as T
    sd = synth_sd16
    prim = T.OctreeOracle(lambda x: O.implicit_forward(sd, x)[:, 0], lambda x: O.implicit_gradient(sd, x)[:, 0, :])
    sec = copy.copy(prim)
    sec.max_iter = 32
    return prim, sec
