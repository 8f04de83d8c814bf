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
    assert ctypes.sizeof(rt.Stats) == 12 * 8
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


def test_header_is_plain_c99(tmp_path):
    """include/stg_rt.h is what a C caller binds: it must compile as strict C99 with a caller that fills every struct."""
    import subprocess
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include "stg_rt.h"
int use(stg_rt_ctx *ctx, double *ranges)
{
  stg_rt_config cfg = { sizeof(stg_rt_config), 0, 50.0, -1, -1, 2, 2, 16, 16, 0, STG_RT_CFG_DEFAULT, 0, 0 };
  stg_rt_fan f = { 1u, 180u, STG_RT_PRED_RANGER, 1, 0, STG_RT_ANGLES_RANGER, 0u, 0.0, 0.0, 0.2, -1.5, 8.0, 3.14 };
  stg_rt_fan_out o = { 0, 0, 0, 0, 0, 0 };
  stg_rt_fid_target t;
  stg_rt_fid_finder ff;
  stg_rt_fiducial fd;
  stg_rt_blob_finder bf;
  stg_rt_blob b;
  stg_rt_ranger_sensor rs;
  stg_rt_ranger_pose rp;
  stg_rt_fan_noise nz;
  stg_rt_stats st;
  int rc;
  o.ranges = ranges;
  (void)t; (void)ff; (void)fd; (void)bf; (void)b; (void)rs; (void)rp; (void)nz; (void)st; (void)cfg;
  rc = stg_rt_fans_submit(ctx, 1u, &f, 0, 0, &o);
  if (!rc) rc = stg_rt_sync(ctx);
  return rc + (int)stg_rt_ranger_rand_advance(0u, 0u);
}
''')
    r = subprocess.run(["gcc", "-std=c99", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"),
                        "-c", str(src), "-o", str(tmp_path / "caller.o")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def test_ranger_rand_advance_makes_the_draws_of_the_beam_loop():
    """stg_rt_ranger_rand_advance leaves libc rand() where ModelRanger::Sensor::Update's beam loop would have left it
    (model_ranger.cc:232-249 with generateGaussianNoise's static spare flag, :181-199), sensor after sensor."""
    import subprocess
    import sys
    code = r'''
import ctypes, sys
sys.path.insert(0, %r)
from stage_b200 import rt
libc = ctypes.CDLL("libc.so.6")
L = rt.lib()
sensors = [(180, [1] * 37 + [0] * 143), (16, [1, 0] * 8), (1, [1]), (1, [1]), (1080, [i %% 3 == 0 for i in range(1080)]), (5, [0] * 5)]
def loop(have_spare):   # the reference's draws, call by call
    for n, hits in sensors:
        for t in range(n):
            libc.rand()                     # simpleNoise() of the heading jitter, :237
            if hits[t]:                     # res.range < range.max, :244
                libc.rand()                 # range_noise * simpleNoise(), :245
                if have_spare:              # generateGaussianNoise(range_noise_const), :246
                    have_spare = False
                else:
                    have_spare = True
                    libc.rand(); libc.rand()
    return [libc.rand() for _ in range(4)]
libc.srand(12345)
want = loop(False)
libc.srand(12345)
for n, hits in sensors:
    L.stg_rt_ranger_rand_advance(n, int(sum(bool(h) for h in hits)))
got = [libc.rand() for _ in range(4)]
assert got == want, (got, want)
print("ok")
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.returncode == 0 and "ok" in r.stdout, r.stderr
