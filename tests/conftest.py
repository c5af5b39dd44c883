import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests", "golden")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build liboetr_b200.so (nvcc cross-compiles without a GPU) if it is not there yet."""
    from oetr_b200 import cabi
    if not os.path.exists(cabi.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    return cabi.LIB_PATH


def load_case(name):
    """(weights dict, feat1, feat2, case tuple, golden npz) for a tests/golden case (CASES or MASK_CASES)."""
    from cases import CASES, MASK_CASES
    from oetr_b200 import weights
    b, fm1, fm2, hw1, hw2, attention, wseed, fseed = {**CASES, **MASK_CASES}[name]
    W = weights.synthetic_hot_path_weights(wseed)
    f1 = weights.synthetic_features(b, *fm1, seed=fseed, tag="feat1")
    f2 = weights.synthetic_features(b, *fm2, seed=fseed, tag="feat2")
    golden = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    return W, f1, f2, {**CASES, **MASK_CASES}[name], golden


def load_masks(name):
    """The synthetic masks of a MASK_CASES case (the generator used them for the committed reference outputs)."""
    from cases import MASK_CASES
    from oetr_b200 import weights
    b, fm1, fm2 = MASK_CASES[name][:3]
    return weights.synthetic_mask(b, *fm1, tag="mask1"), weights.synthetic_mask(b, *fm2, tag="mask2")


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def load_stress_case(name):
    """(weights dict, feat1, feat2, case dict, golden npz) for a tests/golden STRESS_CASES case (trained-like scales)."""
    from cases import STRESS_CASES
    from oetr_b200 import weights
    c = STRESS_CASES[name]
    W = weights.synthetic_hot_path_weights(c["wseed"], ln_gain=c["ln_gain"], head_default_init=c["head_default_init"])
    f1 = weights.synthetic_features(c["batch"], *c["fm1"], seed=c["fseed"], tag="feat1", scale=c["feat_scale"])
    f2 = weights.synthetic_features(c["batch"], *c["fm2"], seed=c["fseed"], tag="feat2", scale=c["feat_scale"])
    return W, f1, f2, c, np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
