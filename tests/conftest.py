import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs /root/reference (build container only)")
    config.addinivalue_line("markers", "slow: minutes of CPU oracle time; runs only with FHC_SLOW=1 in the environment")


def pytest_collection_modifyitems(config, items):
    """Tests marked `gpu` are skipped (not failed) on a machine without a CUDA device, whatever -m says."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if os.environ.get("FHC_SLOW", "0") != "1":
        slow = pytest.mark.skip(reason="slow (minutes of oracle time): set FHC_SLOW=1")
        for item in items:
            if "slow" in item.keywords:
                item.add_marker(slow)
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def lib():
    """The built C-ABI library; built on demand where nvcc exists (the GPU box receives the prebuilt .so)."""
    from fithic_b200 import _capi, build
    if not os.path.exists(_capi.LIB_PATH):
        build.build()
    return _capi.load()


@pytest.fixture(params=["lists", "lists-front-v1", "tile"])
def pval_impl(request, monkeypatch):
    """K3 has two implementations behind fhc_pvalues (work-list pipeline / tile-phased kernel), and the work-list
    pipeline two versions of its front kernel (the first one is what a distance table reaching beyond 2^31 falls back
    to); the library reads FHC_PVAL_IMPL and FHC_PVAL_FRONT on every call, so a module that uses this fixture runs each
    of its tests through all three."""
    impl, _, front = request.param.partition("-front-")
    monkeypatch.setenv("FHC_PVAL_IMPL", impl)
    if front:
        monkeypatch.setenv("FHC_PVAL_FRONT", front)
    else:
        monkeypatch.delenv("FHC_PVAL_FRONT", raising=False)
    return impl
