"""Runs the reference's own tests restated in C++ against the host mirror
(hostcpp/reference_tests.cpp -> hostcpp/poly_b200.hpp -> C ABI -> CUDA)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_reference_tests_through_cpp_mirror(gpu):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hostcpp"), "-s", "CXX=g++"])
    out = subprocess.run([os.path.join(ROOT, "hostcpp", "reference_tests")], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok: reference tests pass" in out.stdout
