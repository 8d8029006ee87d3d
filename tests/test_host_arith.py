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


# ---- carry-free 28-bit-limb arithmetic (fp28.cuh; experimental NTT variant, SS_NTT_RADIX28=1) ----------------
def limbs28(v, n=9):
    out = [(v >> (28 * i)) & (2**28 - 1) for i in range(n - 1)]
    return (ctypes.c_uint32 * n)(*(out + [v >> (28 * (n - 1))]))


def val28(a):
    return sum(int(a[i]) << (28 * i) for i in range(9))


def test_f28_roundtrip_mulc_and_bounds(hc):
    rnd = random.Random(3)
    for v in EDGE + [rnd.randrange(2**256) for _ in range(2000)]:
        assert call(hc.hc_f28_roundtrip, v) == v
    rinv = pow(R, -1, P)
    for _ in range(3000):
        # operand: lazy limbs (each < 2^32), value < 2^256; constant: canonical Montgomery form
        a = (ctypes.c_uint32 * 9)(*([rnd.randrange(2**32) for _ in range(8)] + [rnd.randrange(2**28)]))
        while val28(a) >= 2**256:
            a[8] = rnd.randrange(2**27)
        c = rnd.randrange(P)
        out = (ctypes.c_uint32 * 9)()
        hc.hc_f28_mulc(a, arr(c * R % P), out)
        r = val28(out)
        assert all(out[i] < 2**28 for i in range(8))
        assert r % P == val28(a) * c % P                      # data stays in the R = 2^256 form
        assert r < P + 2**228


def test_f28_weak_reduce_and_biased_sub(hc):
    rnd = random.Random(4)
    for _ in range(5000):
        a = (ctypes.c_uint32 * 9)(*[rnd.randrange(int(2**31.3)) for _ in range(9)])
        if val28(a) >= 2**256:
            a[8] = rnd.randrange(2**27)
        out = (ctypes.c_uint32 * 9)()
        hc.hc_f28_weak_reduce(a, out)
        assert all(out[i] < 2**28 for i in range(8))
        assert val28(out) % P == val28(a) % P and val28(out) < 2**252 + 9 * 2**224
    for dit, bound_b, k in ((0, 2**30, 9), (1, 2**28, 3)):
        for _ in range(3000):
            a = (ctypes.c_uint32 * 9)(*[rnd.randrange(2**30) for _ in range(9)])
            b = (ctypes.c_uint32 * 9)(*[rnd.randrange(bound_b) for _ in range(9)])
            out = (ctypes.c_uint32 * 9)()
            hc.hc_f28_sub(a, b, dit, out)
            assert val28(out) == val28(a) - val28(b) + k * P      # no borrow anywhere: exact integer identity
