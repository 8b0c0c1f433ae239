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


def test_argument_validation_returns_status_and_message():
    """Error conventions of the C ABI (include/step_b200.h): bad arguments are rejected before any CUDA call with a
    negative status and a message retrievable through step_last_error_string(); pure size queries work without a GPU."""
    from step_b200 import lib
    h = lib.load()
    launched = h.step_launch_count()
    # null pointers -> STEP_EINVAL (-1) and a message naming the entry point
    rc = h.step_tc_attention(None, None, None, None, None, 4, 168, 0.0, 0, None)
    assert rc < 0 and b"tc_attention" in h.step_last_error_string()
    rc = h.step_layernorm96_f32(None, None, None, None, 10, None)
    assert rc < 0 and b"layernorm" in h.step_last_error_string()
    rc = h.step_ts_layers_fwd(None, 1, 1, None, 1, None, None, None, 0, 0.0, 0, None)
    assert rc < 0 and b"ts_layers" in h.step_last_error_string()
    # workspace / image size queries are host-only arithmetic
    assert h.step_ts_encoder_workspace_bytes(10, 168) >= 10 * 168 * (96 * 3 + 384) * 4
    q, kv = h.step_tc_attn_image_bytes(3, 168, 0), h.step_tc_attn_image_bytes(3, 168, 1)
    assert q == 3 * 4 * 2 * 6144 and kv == 3 * 4 * 3 * 176 * 16          # 2 row tiles of 128; keys padded to 176
    assert h.step_tc_attn_image_bytes(1, 336, 1) == 4 * 3 * 336 * 16
    assert h.step_tc_attn_image_bytes(3, 168, 2) == 3 * 4 * 169 * 4               # max |k| per head + |q| per query
    assert h.step_tc_seq_image_bytes(2, 207, 168) == 2 * 168 * 12 * 256 * 16
    assert h.step_gwnet_stash_floats(2, 207, 8) > 9 * 2 * 51 * 207 * 32
    assert h.step_launch_count() == launched                               # nothing was launched by the rejected calls
