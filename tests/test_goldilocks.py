"""Goldilocks NTT path (BASELINE config 4).  CPU: the oracle against a quadratic-time big-int DFT and the host build of the
device arithmetic against Python ints.  GPU: ss_ntt / ss_lde with SS_FIELD_GOLDILOCKS against the oracle, every order
combination, sizes 2^0 .. 2^21 (one, two and three passes), and a round trip at 2^24."""
import ctypes
import os
import random

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 2**64 - 2**32 + 1


def test_oracle_matches_definition(oracle):
    rng = np.random.default_rng(1)
    for log_n in (0, 1, 3, 5):
        n = 1 << log_n
        a = rng.integers(0, P, size=(2, n), dtype=np.uint64)
        w = pow(7, (P - 1) // n, P)
        for coset in (False, True):
            h = 7 if coset else 1
            want = [[sum(int(a[c][k]) * pow(h * pow(w, i, P) % P, k, P) for k in range(n)) % P for i in range(n)] for c in range(2)]
            got = oracle.gl_ntt(a, False, coset)
            assert got.tolist() == want
            assert np.array_equal(oracle.gl_ntt(got, True, coset), a)
        coeffs = oracle.gl_ntt(a, True, False)
        N = 2 * n
        wN = pow(7, (P - 1) // N, P)
        want = [[sum(int(coeffs[c][k]) * pow(7 * pow(wN, i, P) % P, k, P) for k in range(n)) % P for i in range(N)] for c in range(2)]
        assert oracle.gl_lde(a, 1).tolist() == want


def test_host_build_of_device_arithmetic():
    path = os.path.join(ROOT, "sandstorm_b200", "_host_check.so")
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build()
    hc = ctypes.CDLL(path)
    for f in (hc.hc_gl_add, hc.hc_gl_sub, hc.hc_gl_mul, hc.hc_gl_reduce128):
        f.restype = ctypes.c_uint64
        f.argtypes = [ctypes.c_uint64, ctypes.c_uint64]
    hc.hc_gl_root.restype = ctypes.c_uint64
    rnd = random.Random(3)
    edge = [0, 1, 2, P - 1, P - 2, 2**32 - 1, 2**32, 2**32 + 1, 2**63, P // 2]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(P), rnd.randrange(P)) for _ in range(20000)]
    for a, b in pairs:
        assert hc.hc_gl_add(a, b) == (a + b) % P
        assert hc.hc_gl_sub(a, b) == (a - b) % P
        assert hc.hc_gl_mul(a, b) == a * b % P
    for lo, hi in [(0, 0), (2**64 - 1, 2**64 - 1), (0, 2**64 - 1), (2**64 - 1, 0), (P, P)] + [(rnd.randrange(2**64), rnd.randrange(2**64)) for _ in range(20000)]:
        assert hc.hc_gl_reduce128(lo, hi) == (lo + (hi << 64)) % P
    for log_n in (1, 8, 32):
        w = hc.hc_gl_root(log_n)
        assert pow(w, 1 << log_n, P) == 1 and pow(w, 1 << (log_n - 1), P) == P - 1


@pytest.fixture(scope="module")
def ss():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import sandstorm_b200

    return sandstorm_b200


def _brev_perm(log_n):
    n = 1 << log_n
    idx = np.arange(n)
    out = np.zeros(n, dtype=np.int64)
    for b in range(log_n):
        out |= ((idx >> b) & 1) << (log_n - 1 - b)
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [0, 1, 2, 5, 11, 12, 13, 14, 16, 21])
def test_gpu_ntt_matches_oracle(ss, oracle, log_n):
    import torch

    from sandstorm_b200 import goldilocks as glk

    rng = np.random.default_rng(40 + log_n)
    n_cols = 3 if log_n <= 16 else 2
    a = rng.integers(0, P, size=(n_cols, 1 << log_n), dtype=np.uint64)
    perm = _brev_perm(log_n)
    for inverse in (False, True):
        for coset in (False, True):
            want = oracle.gl_ntt(a, inverse, coset)
            for in_order, out_order in ((ss.ORDER_NATURAL, ss.ORDER_NATURAL), (ss.ORDER_NATURAL, ss.ORDER_BITREV), (ss.ORDER_BITREV, ss.ORDER_NATURAL)):
                src = a if in_order == ss.ORDER_NATURAL else a[:, perm]
                t = torch.from_numpy(np.ascontiguousarray(src).view(np.int64)).cuda()
                glk.ntt_(t, inverse=inverse, coset=coset, in_order=in_order, out_order=out_order)
                got = t.cpu().numpy().view(np.uint64)
                if out_order == ss.ORDER_BITREV:
                    got = got[:, perm]
                assert np.array_equal(got, want), (log_n, inverse, coset, in_order, out_order)


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,log_blowup", [(0, 1), (3, 1), (10, 2), (12, 1), (15, 3), (19, 1)])
def test_gpu_lde_matches_oracle(ss, oracle, log_n, log_blowup):
    import torch

    from sandstorm_b200 import goldilocks as glk

    rng = np.random.default_rng(70 + log_n)
    a = rng.integers(0, P, size=(2, 1 << log_n), dtype=np.uint64)
    got = glk.lde(torch.from_numpy(a.view(np.int64)).cuda(), log_blowup).cpu().numpy().view(np.uint64)
    assert np.array_equal(got, oracle.gl_lde(a, log_blowup))


@pytest.mark.gpu
@pytest.mark.parametrize("log_n", [24, 25])
def test_gpu_ntt_matches_oracle_two_and_three_passes(ss, oracle, log_n):
    """2^24 = 13 + 11 bits (two passes), 2^25 = 13 + 6 + 6 (three): coset NTT and its inverse against the oracle, one column."""
    import torch

    from sandstorm_b200 import goldilocks as glk

    a = np.random.default_rng(log_n).integers(0, P, size=(1, 1 << log_n), dtype=np.uint64)
    want = oracle.gl_ntt(a, False, True)
    t = torch.from_numpy(a.view(np.int64)).cuda()
    glk.ntt_(t, coset=True)
    assert np.array_equal(t.cpu().numpy().view(np.uint64), want)
    glk.ntt_(t, inverse=True, coset=True)
    assert np.array_equal(t.cpu().numpy().view(np.uint64), a)


@pytest.mark.gpu
def test_gpu_round_trip_at_full_size(ss):
    """2^24 points x 4 columns (the size BASELINE quotes): forward then inverse returns the input, and the transform is
    linear (NTT(a + b) = NTT(a) + NTT(b)) — size-independent properties, no oracle needed."""
    import torch

    from sandstorm_b200 import goldilocks as glk

    g = torch.Generator(device="cuda").manual_seed(9)
    a = torch.randint(0, 2**62, (4, 1 << 24), dtype=torch.int64, device="cuda", generator=g)
    b = torch.randint(0, 2**62, (4, 1 << 24), dtype=torch.int64, device="cuda", generator=g)
    fa = glk.ntt_(a.clone(), coset=True)
    assert torch.equal(glk.ntt_(fa.clone(), inverse=True, coset=True), a)
    fb = glk.ntt_(b.clone(), coset=True)
    s = (a + b)                                   # < 2^63 < p: the plain integer sum is the field sum
    fs = glk.ntt_(s.clone(), coset=True).cpu().numpy().view(np.uint64)
    x, y = fa.cpu().numpy().view(np.uint64), fb.cpu().numpy().view(np.uint64)
    idx = np.random.default_rng(2).integers(0, 1 << 24, 1000)
    for c in range(4):
        for i in idx:
            assert int(fs[c][i]) == (int(x[c][i]) + int(y[c][i])) % P
