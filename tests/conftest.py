import os
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def ctx():
    """One ivlm Context on cuda:0 for the -m gpu tests (fails loudly when the extension is missing)."""
    import torch

    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container")
    from interactvlm_b200 import build as _b

    if _b.needs_build():
        _b.build()
    from interactvlm_b200.ops import Context

    return Context(0)
