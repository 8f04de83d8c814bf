"""The product's port of the host libm's sin/cos (stage_b200/csrc/rt_libm.cuh), built for the host,
against the C library itself, bit for bit. (The device build is checked on the GPU box by
tests/test_gpu_trig.py.) If this host's libm is not the glibc FMA variant the port transcribes, the
test says so instead of failing: strict mode (host-supplied trig) is then the exact path."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    so = tmp_path_factory.mktemp("libmport") / "libm_port_check.so"
    subprocess.run(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-fPIC", "-shared",
                    os.path.join(HERE, "libm_port_check.cc"), "-o", str(so), "-lm"], check=True)
    L = ctypes.CDLL(str(so))
    for f in (L.port_sincos, L.libm_sincos):
        f.argtypes = [ctypes.c_uint64, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    return L


def both(L, x):
    x = np.ascontiguousarray(x, np.float64)
    out = [np.empty_like(x) for _ in range(4)]
    L.port_sincos(len(x), x.ctypes.data, out[0].ctypes.data, out[1].ctypes.data)
    L.libm_sincos(len(x), x.ctypes.data, out[2].ctypes.data, out[3].ctypes.data)
    return [o.view(np.uint64) for o in out]


def host_is_glibc_fma():
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return " fma " in flags and " avx2 " in flags and os.path.exists("/lib/x86_64-linux-gnu/libm.so.6")


def test_port_equals_libm_bit_for_bit(lib):
    rng = np.random.RandomState(1)
    xs = [rng.uniform(-7.0, 7.0, 4_000_000),                                   # every heading a fan can have
          np.float32(rng.uniform(-7.0, 7.0, 2_000_000)).astype(np.float64),      # ranger headings are float-valued
          rng.uniform(-0.13, 0.13, 500_000), rng.uniform(0.85, 0.86, 200_000), rng.uniform(2.42, 2.43, 200_000),
          rng.uniform(-1000.0, 1000.0, 1_000_000), 10.0 ** rng.uniform(-12, 8, 1_000_000) * rng.choice([-1, 1], 1_000_000),
          np.array([1e-12, -1e-12, 0.126, -0.126, 0.855469, 2.426265, np.pi, -np.pi, np.pi / 2, -np.pi / 2, 3 * np.pi / 2,
                    2 * np.pi, 0.5, 1.0, 2.0, 4.0, 105414349.0, 2.0 ** -26, 2.0 ** -27, 2.0 ** -28])]
    x = np.concatenate(xs)
    x = x[np.abs(x) < 105414350.0]
    ps, pc, ls, lc = both(lib, x)
    bad_s, bad_c = int(np.count_nonzero(ps != ls)), int(np.count_nonzero(pc != lc))
    if not host_is_glibc_fma() and (bad_s or bad_c):
        pytest.skip("host libm is not glibc's x86-64 FMA build: %d sin / %d cos of %d differ" % (bad_s, bad_c, len(x)))
    assert bad_s == 0 and bad_c == 0, "sin: %d, cos: %d of %d arguments differ from libm" % (bad_s, bad_c, len(x))
