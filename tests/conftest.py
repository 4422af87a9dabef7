import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


def _gpu_unavailable_reason():
    """GPU tests need a CUDA device AND the in-tree C-ABI library; on a CPU box (or before `make`) they are skipped, not failed."""
    try:
        import torch

        if not torch.cuda.is_available():
            return "needs a B200 (torch.cuda.is_available() is False)"
    except Exception as e:  # noqa: BLE001
        return f"torch unavailable: {e}"
    lib = os.path.join(ROOT, "ant-multi-modal-framework_b200", "libb200mm.so")
    if not os.path.isfile(lib):
        return "libb200mm.so is not built (python -c 'import __graft_entry__ as g; g.build()')"
    return None


def pytest_collection_modifyitems(config, items):
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
