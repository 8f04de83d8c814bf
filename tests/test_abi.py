"""The drop-in boundary: libstg_rt.so loads and exports exactly what include/stg_rt.h declares.
No compute calls here (no GPU in the CPU test tier)."""
import ctypes
import os
import re

import pytest

from stage_b200 import build, rt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "stg_rt.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(stg_rt_[a-z_0-9]+)\s*\(", src)))


def test_library_builds_and_exports_every_declared_symbol():
    path = build.build()
    L = ctypes.CDLL(path)
    names = declared_functions()
    assert len(names) >= 25
    for n in names:
        assert hasattr(L, n), "libstg_rt.so does not export %s" % n
    assert sorted(rt.EXPORTED_SYMBOLS) == names


def test_struct_layouts_match_the_header():
    assert rt.FAN_DTYPE.itemsize == 64
    assert ctypes.sizeof(rt.Config) == 56
    assert ctypes.sizeof(rt.FanOut) == 6 * 8
    assert ctypes.sizeof(rt.Stats) == 11 * 8
    assert rt.lib().stg_rt_abi_version() == 1


def test_no_cpu_fallback_create_fails_loudly_without_a_gpu():
    try:
        import torch
        if torch.cuda.is_available():
            pytest.skip("a GPU is present")
    except ImportError:
        pass
    with pytest.raises(rt.StgRtError) as e:
        rt.Context(50.0, -1, -1, 2, 2)
    assert e.value.code == -6  # STG_RT_ENODEV
    assert "no CPU path" in str(e.value)


def test_bad_config_is_rejected():
    L = rt.lib()
    cfg = rt.Config(12, 0, 50.0, 0, 0, 1, 1, 16, 16, 0, 0, 0, 0)  # wrong struct_size
    h = ctypes.c_void_p()
    assert L.stg_rt_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b"struct_size" in L.stg_rt_last_error(None)


def test_product_does_not_touch_the_oracle():
    """The product path must never import, link or execute anything under oracle/."""
    pkg = os.path.join(ROOT, "stage_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".h", ".cpp", ".cc")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.lower(), "%s mentions the oracle" % f
    import subprocess
    out = subprocess.run(["ldd", build.LIB], capture_output=True, text=True).stdout
    assert "t1_oracle" not in out and "stage_ref" not in out
