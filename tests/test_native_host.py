"""Host-compiled checks of device-side logic that does not need a GPU (nvcc builds plain host code here)."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not on PATH")
def test_top8_selection_variants_agree_with_a_plain_sort(tmp_path):
    """pin_slam_b200/csrc/knn_select.cuh: the register-only selection (product build) and the scratch variant
    (PINB_K1_SMEM_SELECT) return the 8 smallest candidates in ascending order, bit-identically, for 200k random
    candidate streams with and without ties (tests/native/select_equiv.cu)."""
    exe = tmp_path / "select_equiv"
    subprocess.check_call(["nvcc", "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-o", str(exe),
                           os.path.join(ROOT, "tests", "native", "select_equiv.cu")])
    out = subprocess.check_output([str(exe)]).decode()
    assert out.startswith("ok 200000"), out
