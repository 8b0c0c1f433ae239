"""CPU: the C-ABI shared library builds, loads, and exports every symbol include/step_b200.h declares
(no compute calls - there is no GPU here)."""
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "step_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(step_[a-z0-9_]+)\s*\(", text)))


def test_library_builds_and_exports_declared_symbols():
    from step_b200 import build, lib
    build.build()
    handle = lib.load()
    names = declared_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(handle, n), f"{n} declared in step_b200.h but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature"
    assert sorted(lib.SIGNATURES) == names
    assert handle.step_abi_version() == lib.ABI_VERSION
    assert handle.step_last_error_string() is not None


def test_sass_is_sm100a():
    import subprocess
    from step_b200 import build
    out = subprocess.run(["cuobjdump", "-lelf", build.build()], capture_output=True, text=True).stdout
    assert "sm_100a" in out
