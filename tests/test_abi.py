"""The C-ABI shared library: builds for sm_100a, loads on a box without a GPU, exports exactly the
symbols include/fitnoise_b200.h declares, and refuses to compute without a device."""
import ctypes
import os
import re

from conftest import ROOT
from fitnoise_b200 import _native, build


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "fitnoise_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fn_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    assert os.path.exists(path)
    lib = ctypes.CDLL(path)
    names = declared_symbols()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), "missing export %s" % name
    assert sorted(_native.EXPORTS) == names


def test_version_and_device_count():
    lib = _native.load_library()
    assert b"sm_100a" in lib.fn_version()
    assert lib.fn_device_count() >= 0


def test_family_struct_layout():
    assert ctypes.sizeof(_native.Family) == 40
    assert ctypes.sizeof(_native.PrepInfo) == 40
