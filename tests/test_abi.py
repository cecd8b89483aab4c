"""CPU tier: the C-ABI library loads, exports every symbol include/poly_b200.h declares,
and fails LOUDLY (no CPU fallback) when no GPU is usable.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from poly_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "poly_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pg_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    lib = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/poly_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), set(names) ^ set(_lib.EXPORTS)
    assert lib.pg_version() >= 100


def test_product_never_imports_oracle():
    """Only tests/, smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "poly_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", ".hpp")):
                assert "oracle" not in open(os.path.join(dp, f), errors="replace").read().lower().replace("cpu oracle", ""), f
    for f in ("poly_b200.h",):
        assert "oracle" not in open(os.path.join(ROOT, "include", f)).read().lower()


def _no_gpu():
    n = C.c_int(0)
    _lib.lib().pg_device_count(C.byref(n))
    return n.value == 0


@pytest.mark.skipif(not _no_gpu(), reason="this check is for the GPU-less tier")
def test_no_cpu_fallback_fails_loudly():
    from poly_b200 import align, mash, primers

    with pytest.raises(_lib.PolyError) as e:
        mash.New(21, 1000).Sketch("ACGT" * 50)
    assert e.value.code == _lib.PG_ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(_lib.PolyError):
        primers.MeltingTemp("GTAAAACGACGGCCAGT")
    with pytest.raises(_lib.PolyError):
        align.SmithWatermanScore("GATTACA", "GCATGCU", align.NewScoring(None, -1))
    with pytest.raises(_lib.PolyError):
        mash.sketch_uniform(np.zeros(150 * 32, np.uint8), 32, 150, 21, 1000)


def test_host_mirror_marshalling():
    """Pure host logic: flatten([]string) and the alphabet -> byte LUT flattening."""
    from poly_b200 import align, mash

    bases, off = mash.flatten(["ACG", "", b"TT", np.frombuffer(b"N", np.uint8)])
    assert bytes(bases) == b"ACGTTN" and off.tolist() == [0, 3, 3, 5, 6]
    a = align.NewAlphabet(["-", "A", "C", "G", "T"])
    lut = a.byte_lut()
    assert lut[ord("A")] == 1 and lut[ord("-")] == 0 and lut[ord("N")] == -1 and (lut >= 0).sum() == 5
    with pytest.raises(align.AlphabetError, match="Symbol N not in alphabet"):
        a.Encode("N")
    assert align.Default.Score("A", "A") == 1 and align.Default.Score("A", "Z") == -1
    m = align.NewSubstitutionMatrix(a, a, align.NUC_4)
    assert m.Score("A", "A") == 5 and m.Score("C", "T") == -4 and m.Score("-", "-") == 0  # matrix_test.go:11-48
    with pytest.raises(ValueError):
        align.NewSubstitutionMatrix(a, a, [[1, 2], [3, 4]])
    with pytest.raises(IndexError):
        mash.New(21, -1)
