"""deplex_b200/csrc/cr_math.cuh on the CPU: the double-double sin / cos / atan2 that region_grow.cu repair_axis_cell uses for
the cells whose histogram bin hangs on the last ulp of the eigen-solver's trigonometry.  They are meant to be correctly
rounded; glibc (the reference's libm) is correctly rounded for all but ~0.2 % of arguments, so the two must agree almost
everywhere and never be more than one ulp apart."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def crm():
    out = os.path.join(tempfile.mkdtemp(prefix="crm_"), "libcrm.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", out,
                           os.path.join(HERE, "native", "crm_harness.cpp")])
    lib = C.CDLL(out)
    dp = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
    lib.crm_sincos.argtypes = [dp, dp, dp, C.c_long]
    lib.crm_atan2.argtypes = [dp, dp, dp, dp, C.c_long]
    return lib


def _ulps(a, b):
    return np.abs(a.view(np.int64) - b.view(np.int64))


def test_sincos_matches_glibc_to_the_last_bit_almost_everywhere(crm):
    rng = np.random.default_rng(3)
    a = np.concatenate([rng.uniform(0, np.pi / 3, 400_000), rng.uniform(0, 1e-3, 50_000),
                        np.pi / 3 - rng.uniform(0, 1e-6, 50_000), rng.uniform(0, 3.2, 100_000)])
    s, c = np.empty_like(a), np.empty_like(a)
    crm.crm_sincos(a, s, c, a.size)
    import math
    # (math.*, not numpy: numpy's vectorised loops are not glibc's functions)
    for got, ref in ((s, np.array([math.sin(v) for v in a])), (c, np.array([math.cos(v) for v in a]))):
        d = _ulps(got, ref)
        assert d.max() <= 1
        assert (d != 0).mean() < 5e-3      # glibc's own misroundings (error bound 0.55 ulp)


def test_atan2_from_a_perturbed_start_matches_glibc(crm):
    rng = np.random.default_rng(4)
    n = 400_000
    y = rng.uniform(0, 1, n) * np.where(rng.random(n) < 0.2, 1e-9, 1.0)
    x = rng.uniform(-1, 1, n) * np.where(rng.random(n) < 0.1, 1e-9, 1.0)
    import math
    ref = np.array([math.atan2(p, q) for p, q in zip(y, x)])  # glibc's, not numpy's vectorised loop (several ulp)
    a0 = ref.copy()
    for k in range(1, 3):                   # start 0, 1 or 2 ulp off, either side
        up = rng.integers(-1, 2, n)
        a0 = np.where(up > 0, np.nextafter(a0, 10.0), np.where(up < 0, np.nextafter(a0, -1.0), a0))
    out = np.empty_like(ref)
    crm.crm_atan2(y, x, a0, out, n)
    d = _ulps(out, ref)
    assert d.max() <= 1
    assert (d != 0).mean() < 5e-3


def test_axis_aware_bin_angles_equal_glibc():
    """normal_bins.cuh atan2_for_bins / acos_for_bins, restated with Python floats (IEEE fp64, the same operations in the
    same order): pi - tiny and pi/2 - tiny come out exactly as glibc's atan2 / acos give them."""
    import math
    near, pi_lo, p2_lo = 2.0 ** -20, 1.2246467991473532e-16, 6.123233995736766e-17

    def small_atan(t):
        return t - (t * t * t) / 3.0

    def at2(s, c):
        a_s, a_c = abs(s), abs(c)
        if c < 0.0 and a_s < near * a_c:
            r = math.pi + (pi_lo - small_atan(a_s / a_c))
            return -r if math.copysign(1.0, s) < 0 else r
        if a_s > 0.0 and a_c < near * a_s:
            r = math.pi / 2 + (p2_lo - small_atan(c / a_s))
            return -r if s < 0 else r
        return math.atan2(s, c)

    def ac(x):
        if abs(x) < near:
            return math.pi / 2 + (p2_lo - (x + (x * x * x) / 6.0))
        return math.acos(x)

    rng = np.random.default_rng(5)
    for _ in range(200_000):
        e = 10.0 ** rng.uniform(-17, -6.05) * rng.choice([-1.0, 1.0])
        big = rng.uniform(0.1, 1.0) * rng.choice([-1.0, 1.0])
        assert at2(e, big) == math.atan2(e, big)
        assert at2(big, e) == math.atan2(big, e)
        assert ac(e) == math.acos(e)
    assert at2(0.0, -1.0) == math.pi and at2(-0.0, -1.0) == -math.pi
