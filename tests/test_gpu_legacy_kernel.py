"""The unstructured Gram kernel (gram_fused_kernel, round 1) stays the fallback for trees the structured kernel does not serve
and the A/B reference for it: the same oracle comparisons, run in a child process with SYSID_GRAM_LEGACY=1 (the switch is read
once per process)."""
import os
import subprocess
import sys

import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_unstructured_kernel_passes_the_same_gram_parity_tests():
    env = dict(os.environ, SYSID_GRAM_LEGACY="1", SYSID_DEBUG_KERNEL="1")
    res = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-x", "-q", "-m", "gpu", "-s",
                          "-k", "fused_gram_vs_golden or gram_edge_cases or bootstrap_resamples or full_size_properties"],
                         cwd=ROOT, env=env, capture_output=True, text=True, timeout=900)
    out = res.stdout + res.stderr
    assert res.returncode == 0, out[-3000:]
    assert "gram kernel: unstructured" in out and "gram kernel: structured" not in out, out[-2000:]
