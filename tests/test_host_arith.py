"""CPU-only: the __host__ __device__ field arithmetic of the CUDA kernels (PTX carry primitives
emulated, sandstorm_b200/csrc/arith.cuh) checked against Python big ints, including the lazy-domain
bounds the kernels rely on."""
import ctypes
import os
import random

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**251 + 17 * 2**192 + 1
R = 2**256
M = 2**252 + 2**224


@pytest.fixture(scope="module")
def hc():
    path = os.path.join(ROOT, "sandstorm_b200", "_host_check.so")
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build()
    return ctypes.CDLL(path)


def arr(v):
    return (ctypes.c_uint32 * 8)(*[(v >> (32 * i)) & 0xFFFFFFFF for i in range(8)])


def val(a):
    return sum(int(a[i]) << (32 * i) for i in range(8))


def call(fn, *xs):
    out = (ctypes.c_uint32 * 8)()
    fn(*[arr(x) for x in xs], out)
    return val(out)


EDGE = [0, 1, P - 1, P, P + 1, 2 * P - 1, 2 * P, M - 1, 2**252, 2**251, 2**224 - 1, 2**192, 2**64 - 1, 2**32 - 1, 2**253 - 1, 4 * P - 1, 3 * P]


def test_montgomery_mul_is_exact_and_lazy_bounded(hc):
    rnd = random.Random(1)
    rinv = pow(R, -1, P)
    cases = [(a, b) for a in EDGE for b in EDGE] + [(rnd.randrange(2**255), rnd.randrange(2**255)) for _ in range(5000)]
    for a, b in cases:
        out = (ctypes.c_uint32 * 8)()
        hc.hc_fp_mul(arr(a), arr(b), out, 0)
        r = val(out)
        assert r % P == a * b * rinv % P
        assert a * b // R <= r <= a * b // R + P
        hc.hc_fp_mul(arr(a % P), arr(b % P), out, 1)
        assert val(out) == (a % P) * (b % P) * rinv % P


def test_lazy_add_sub_canon(hc):
    rnd = random.Random(2)
    B = 2**253
    for _ in range(5000):
        a = rnd.choice(EDGE + [rnd.randrange(B)] * 3) % B
        b = rnd.choice(EDGE + [rnd.randrange(B)] * 3) % B
        r = call(hc.hc_fp_add, a, b)
        assert r % P == (a + b) % P and r < M
        r = call(hc.hc_fp_sub, a, b)
        assert r % P == (a - b) % P and r < M
        assert call(hc.hc_fp_sub4p, a, b) == a - b + 4 * P
        t, x = rnd.randrange(2 * P), rnd.randrange(6 * P)
        assert call(hc.hc_fp_sub2p, x, t) == x - t + 2 * P
        y = rnd.randrange(8 * P)
        r = call(hc.hc_fp_reduce8p, y)
        assert r % P == y % P and r < M
        z = rnd.randrange(4 * P + 2**224)
        assert call(hc.hc_fp_canon, z) == z % P


def test_inverse_and_constants(hc):
    for v in (1, 2, 3, P - 1, 123456789123456789):
        assert call(hc.hc_fp_inv, v * R % P) == pow(v, -1, P) * R % P
    out = (ctypes.c_uint32 * 8)()
    hc.hc_fp_from_u32(12345, out)
    assert val(out) == 12345 * R % P


