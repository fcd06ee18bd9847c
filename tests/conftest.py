import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "slow: parity at the benchmark shapes (minutes of oracle CPU time)")


def _cuda_device_visible():
    """True when the product library can open a CUDA context (xb_ctx_create != XB_ERR_NO_DEVICE)."""
    try:
        import xped_b200 as xb
        ctx = xb.Context([("U0", 0, 0)])
        ctx.close()
        return True
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a machine without a GPU: tests marked gpu are skipped instead of failing (the product has no CPU fallback)."""
    gpu_items = [it for it in items if it.get_closest_marker("gpu")]
    if not gpu_items or _cuda_device_visible():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible (xb_ctx_create: XB_ERR_NO_DEVICE)")
    for it in gpu_items:
        it.add_marker(skip)
