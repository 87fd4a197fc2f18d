"""CPU: the C-ABI library loads and exports every symbol include/psl_frontend.h declares."""
import ctypes
import os
import re

from conftest import ROOT


def _declared():
    hdr = open(os.path.join(ROOT, "include", "psl_frontend.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(psl_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    from psl_slam_b200 import _lib
    assert os.path.exists(_lib.SO_PATH), "build first: python -c 'import __graft_entry__ as g; g.build()'"
    lib = ctypes.CDLL(_lib.SO_PATH)
    names = _declared()
    assert len(names) >= 10
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_lib.EXPORTS) == names


def test_no_gpu_is_a_loud_error():
    """Without a B200 psl_create must fail (no CPU fallback)."""
    import torch

    from psl_slam_b200 import ORBextractor, PslError
    if torch.cuda.is_available():
        return
    try:
        ORBextractor()
    except PslError as e:
        assert e.code == -2
    else:
        raise AssertionError("psl_create succeeded without a GPU")


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "psl_slam_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"(import|from|include|CDLL).*oracle", src), os.path.join(dp, f)
