import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a B200 (run with -m gpu on the GPU box)')


@pytest.fixture(scope='session')
def golden_dir():
    return GOLDEN


@pytest.fixture(scope='session')
def libnbc():
    """Build (if stale) and load libnbc.so."""
    from neuralbarkcalculator_b200 import _lib, build
    if build.needs_build():
        build.build()
    return _lib.load()


@pytest.fixture(scope='session')
def cuda_device(libnbc):
    import torch
    if not torch.cuda.is_available():
        pytest.fail('GPU test selected but no CUDA device is visible (there is no CPU fallback to test)')
    from neuralbarkcalculator_b200 import _lib
    _lib.require_device(0)
    return torch.device('cuda:0')


@pytest.fixture(scope='session')
def synthetic_sd():
    """Seed-0 synthetic state_dict with the calibrated head stored in the golden fixture (bit-reproducible)."""
    import numpy as np
    from oracle import model as omodel
    g = np.load(os.path.join(GOLDEN, 'model_small.npz'))
    return omodel.synthetic_state_dict(seed=int(g['state_dict_seed']), head=(g['head_w'], g['head_b']))


@pytest.fixture(scope='session')
def trained_like_sd():
    """Synthetic state_dict conditioned like a TRAINED ResNet (the last BatchNorm of every bottleneck has gain 0.1, cf.
    torchvision's zero_init_residual; the classifier is three random directions on the standardised head features, the
    noise gain of an ordinary linear layer) with unit-scale logits (overall std ~1).  The network the north_star
    floating-point bar is gated on (tests/test_gpu_model.py)."""
    from oracle import model as omodel
    return omodel.synthetic_state_dict(seed=3, branch_gain=0.1, calibration='features', logit_std=0.6,
                                       class_bias=(0.8, 0.1, -1.1))
