import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """The CUDA library and the oracle are built in-tree before anything imports them
    (prebuilt files that travelled with the snapshot are reused as-is)."""
    pkg_dir = os.path.join(ROOT, "k8s-gpu-hpa_b200")
    need = [os.path.join(pkg_dir, f) for f in ("libb200va.so", "vectorAdd")]
    need.append(os.path.join(ROOT, "oracle", "liboracle_vadd.so"))
    if not all(os.path.exists(p) for p in need):
        import __graft_entry__

        __graft_entry__.build()


def has_gpu() -> bool:
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
