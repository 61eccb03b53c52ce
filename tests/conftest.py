import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "critic-vae_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)


def _limit_cpu_threads():
    # the CPU oracle must not oversubscribe a shared host (GPU boxes expose many cores to few tenants)
    try:
        import torch
        torch.set_num_threads(max(1, min(8, os.cpu_count() or 1)))
    except Exception:
        pass


def pytest_configure(config):
    _limit_cpu_threads()
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def critic_state():
    import torch
    path = os.path.join(ROOT, "critic-vae_b200", "saved-networks",
                        "critic-rewidx=1-cepochs=15-datamode=trunk-datasize=99999-shift=12-chfak=1-dropout=0.3.pt")
    return torch.load(path, map_location="cpu")
