import os
import sys
import hashlib

import numpy as np
import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLD = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def sd_digest(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(v.detach().cpu().numpy().tobytes())
    return h.hexdigest()


def T(a):
    return torch.from_numpy(np.ascontiguousarray(a))


@pytest.fixture(scope="session")
def synth():
    from ratio_guided_multimodal_fm_b200 import synthetic
    cache = {}

    def get(kind, seed):
        key = (kind, int(seed))
        if key not in cache:
            cache[key] = synthetic.state_dict(kind, int(seed))
        return cache[key]
    return get
