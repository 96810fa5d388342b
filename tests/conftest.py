import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


# The training-step tests build VGGLoss; the pretrained VGG19 checkpoint cannot be downloaded offline and the product
# refuses to fall back to random weights silently, so the test-suite opts in explicitly (test_host_cpu.py checks the
# refusal itself).
os.environ.setdefault("JPDSE_VGG_RANDOM", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="session")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.fail("GPU test selected but no CUDA device is visible (there is no CPU fallback)")
    return torch.device("cuda", 0)
