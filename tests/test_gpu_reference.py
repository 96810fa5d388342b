"""GPU: the real reference (baseline/_ref, unmodified) on the B200 beside the drop-in."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "ctu", "__init__.py")),
                    reason="reference not installed in baseline/_ref (tools/install_reference.sh)")
@pytest.mark.parametrize("H,W", [(128, 256), (512, 1024)])
def test_reference_trainer_on_cuda_vs_drop_in(cuda, H, W):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_gpu_check.py"), str(H), str(W)],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    sys.stdout.write(r.stdout[-1500:])
    assert r.returncode == 0 and "REFERENCE_GPU_CHECK_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]


@pytest.mark.skipif(not os.path.isfile(os.path.join(ROOT, "baseline", "_ref", "ctu", "__init__.py")),
                    reason="reference not installed in baseline/_ref (tools/install_reference.sh)")
def test_reference_trainer_train_step_vs_mirror(cuda):
    """One whole training step of the unmodified reference trainer on CUDA fp32 against the mirror trainer on the sm_100a
    kernels, same initial weights of netG / netD / VGG19 and the same batch: the six losses and the first Adam update."""
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_train_check.py"), "128", "256"],
                       capture_output=True, text=True, timeout=900, env=dict(os.environ, PYTHONDONTWRITEBYTECODE="1"))
    sys.stdout.write(r.stdout[-2500:])
    assert r.returncode == 0 and "REFERENCE_TRAIN_CHECK_OK" in r.stdout, r.stdout[-2500:] + r.stderr[-3000:]
