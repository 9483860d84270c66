import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on a B200 with `-m gpu`)")
    config.addinivalue_line("markers", "multigpu: needs >= 2 CUDA devices on one box (run with `gpurun --gpus 2 -- python -m pytest tests -m multigpu`); "
                                       "kept out of `-m gpu` so the 1-GPU run has nothing to skip")


def _n_cuda():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


def pytest_collection_modifyitems(config, items):
    n = _n_cuda()
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for it in items:
        if "gpu" in it.keywords and n < 1:
            it.add_marker(skip)
    # tests for >= 2 devices are deselected (not skipped) where there are fewer: run them with `gpurun --gpus 2 ... -m multigpu`
    if n < 2:
        drop = [it for it in items if "multigpu" in it.keywords]
        if drop:
            items[:] = [it for it in items if "multigpu" not in it.keywords]
            config.hook.pytest_deselected(items=drop)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library (built here with nvcc if missing; never a fallback)."""
    import flashattn_b200._cabi as cabi
    return cabi.load(build_if_missing=True)
