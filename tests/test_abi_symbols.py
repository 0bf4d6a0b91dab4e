"""The C-ABI library loads without a GPU and exports every symbol the headers declare."""
import ctypes
import glob
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        txt = open(h).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        names.update(re.findall(r"\b(sb200_[a-z0-9_]+)\s*\(", txt))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    from stract_b200 import _lib
    L = ctypes.CDLL(_lib._SO)
    decl = _declared()
    assert len(decl) >= 15
    missing = [n for n in decl if not hasattr(L, n)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"


def test_error_path_without_gpu_is_clean():
    from stract_b200 import lib
    L = lib()
    assert L.sb200_version().startswith(b"stract_b200")
    # NULL handle -> error code + message, no crash
    assert L.sb200_hyperball_reset(None) != 0
    assert b"NULL" in L.sb200_last_error()
