"""Runs the reference's own tests restated in C++ against the host mirror
(hostcpp/reference_tests.cpp -> hostcpp/poly_b200.hpp -> C ABI -> CUDA)."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def test_reference_tests_through_cpp_mirror(gpu):
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "hostcpp"), "-s", "CXX=g++"])
    import tempfile

    from poly_b200 import sketchfile

    with tempfile.TemporaryDirectory() as td:
        path = os.path.join(td, "set.pgsketch")
        out = subprocess.run([os.path.join(ROOT, "hostcpp", "reference_tests"), path], capture_output=True, text=True, timeout=300)
        assert out.returncode == 0, out.stdout + out.stderr
        assert "ok: reference tests pass" in out.stdout
        # the PGSKETCH file the C++ mirror wrote is read by the Python implementation and written back identically
        sk, count, k, s = sketchfile.load(path)
        assert (k, s) == (21, 64) and len(count) == 41 and count[-1] == 3 and list(sk[-1, :3]) == [1, 2, 3] and (count[:-1] == 64).all()
        again = os.path.join(td, "again.pgsketch")
        sketchfile.save(again, sk, count, k, s)
        assert open(again, "rb").read() == open(path, "rb").read()
