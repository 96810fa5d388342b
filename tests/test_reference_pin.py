"""CPU, only where /root/reference exists (the build container): re-runs the pin of the oracle against the
imported, unmodified reference. On the GPU box the reference is absent and these tests skip; the golden
vectors it produced are checked by test_oracle_golden.py everywhere."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ctu")), reason="reference tree not mounted here")
def test_oracle_is_bit_identical_to_reference(tmp_path):
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1", JPDSE_GOLDEN_DIR=str(tmp_path))
    r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "pin_against_reference.py")], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "preprocess: oracle == reference" in r.stdout
    assert "generator: oracle == reference" in r.stdout
    assert "quantisers: oracle == reference" in r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "ctu")), reason="reference tree not mounted here")
def test_install_into_reference_builds_our_generator_behind_the_reference_trainer():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "reference_install_check.py"), REF, ROOT], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "loaded the reference checkpoint" in r.stdout


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "datasets", "cityscapes_test_CVPR20_1024")), reason="reference dataset not mounted here")
def test_compact_host_loader_equals_the_reference_loader():
    """jpd-se_b200/ctu/data against the unmodified reference DataLoader on its bundled Cityscapes files."""
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "loader_pin_check.py"), REF, ROOT], env=env,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "LOADER_PIN_OK" in r.stdout, r.stdout[-1500:] + r.stderr[-3000:]
