"""GPU parity: FRI folding and out-of-domain polynomial evaluation vs big-int restatements."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ss():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import sandstorm_b200

    return sandstorm_b200


def fold_reference(oracle, evals_int, log_n, log_fold, alpha, h):
    """out[i] = sum_j alpha^j F_j(y_i): interpolate f on h*<w_N>, regroup coefficients, evaluate on h^F*<w_{N/F}>."""
    P = oracle.P
    N, F = 1 << log_n, 1 << log_fold
    coeffs = oracle.from_mont(oracle.ntt(oracle.to_mont(evals_int)[None], inverse=True)[0])     # of f(h x)
    hinv = pow(h, -1, P)
    coeffs = [c * pow(hinv, k, P) % P for k, c in enumerate(coeffs)]                              # of f
    m = N // F
    g = [sum(pow(alpha, j, P) * coeffs[F * k + j] for j in range(F)) % P for k in range(m)]      # G(y) = sum alpha^j F_j(y)
    hf = pow(h, F, P)
    g = [c * pow(hf, k, P) % P for k, c in enumerate(g)]
    return oracle.from_mont(oracle.ntt(oracle.to_mont(g)[None])[0])


@pytest.mark.parametrize("log_n,log_fold", [(3, 1), (4, 2), (6, 3), (8, 4), (10, 3), (13, 3), (16, 3)])
def test_fri_fold_matches_definition(ss, oracle, log_n, log_fold):
    import torch

    from sandstorm_b200.matrix import fri_fold

    rng = np.random.default_rng(50 + log_n)
    ev = oracle.random_felts(rng, 1 << log_n)
    ev_int = oracle.from_mont(ev)
    alpha = int.from_bytes(rng.bytes(31), "big")
    h = 3
    want = fold_reference(oracle, ev_int, log_n, log_fold, alpha, h)
    d = torch.from_numpy(ev.view(np.int64)).cuda()
    got = fri_fold(d, log_fold, oracle.to_mont([alpha])[0], oracle.to_mont([h])[0])
    torch.cuda.synchronize()
    assert oracle.from_mont(got.cpu().numpy().view(np.uint64)) == want
    got8 = fri_fold(d, log_fold, oracle.to_mont([alpha])[0], oracle.to_mont([h])[0], starkware_scale=True)
    assert oracle.from_mont(got8.cpu().numpy().view(np.uint64)) == [v * (1 << log_fold) % oracle.P for v in want]


def test_fri_fold_equals_repeated_binary_folds(ss, oracle):
    """Fold by 8 with alpha == three folds by 2 with alpha, alpha^2, alpha^4 (the StarkWare construction)."""
    import torch

    from sandstorm_b200.matrix import fri_fold

    rng = np.random.default_rng(60)
    P = oracle.P
    ev = oracle.random_felts(rng, 1 << 12)
    alpha = int.from_bytes(rng.bytes(31), "big")
    d = torch.from_numpy(ev.view(np.int64)).cuda()
    mont = lambda v: oracle.to_mont([v])[0]
    by8 = fri_fold(d, 3, mont(alpha), mont(3))
    step, h = d, 3
    for k in range(3):
        step = fri_fold(step, 1, mont(pow(alpha, 1 << k, P)), mont(h))
        h = h * h % P
    torch.cuda.synchronize()
    assert np.array_equal(by8.cpu().numpy(), step.cpu().numpy())


def test_fri_layer_commit_is_a_strided_view(ss, oracle):
    """Row i of the FRI layer matrix is (e[i], e[i+N/8], ...): committing the evaluation buffer with
    col_stride = N/8 equals the oracle tree over the explicitly transposed matrix."""
    import ctypes

    import torch

    rng = np.random.default_rng(61)
    N = 1 << 10
    ev = oracle.random_felts(rng, N)
    d = torch.from_numpy(ev.view(np.int64)).cuda()
    c = ss.default_context()
    handle = ctypes.c_void_p()
    c.check(c.lib.ss_merkle_build(c.handle, ss.TREE_KECCAK_M20, 0, ctypes.c_void_p(d.data_ptr()), N // 8, 8, 7, ss.ORDER_NATURAL,
                                  ctypes.byref(handle), None))
    root = (ctypes.c_uint8 * 32)()
    c.check(c.lib.ss_merkle_root(c.handle, handle, root))
    c.lib.ss_tree_free(handle)
    cols = np.ascontiguousarray(ev.reshape(8, N // 8, 4))
    assert bytes(root) == oracle.merkle_build(oracle.TREE_KECCAK_M20, cols)[2]


@pytest.mark.parametrize("log_n", [0, 1, 3, 8, 11, 12, 14, 22])
def test_poly_eval_matches_horner(ss, oracle, log_n):
    from sandstorm_b200.matrix import poly_eval

    rng = np.random.default_rng(70 + log_n)
    P = oracle.P
    n_cols = 3 if log_n < 20 else 1
    cols = oracle.random_felts(rng, n_cols, 1 << log_n)
    lde, coeffs = ss.Matrix.from_numpy(cols).lde(1, keep_coeffs=True)
    pts = [int.from_bytes(rng.bytes(31), "big") for _ in range(4)]
    which = [0, n_cols - 1, 0, n_cols // 2]
    got = oracle.from_mont(poly_eval(coeffs, which, oracle.to_mont(pts)))
    if log_n <= 14:
        plain = oracle.ntt(cols, inverse=True)
        for e, (c, z) in enumerate(zip(which, pts)):
            acc = 0
            for coef in reversed(oracle.from_mont(plain[c])):
                acc = (acc * z + coef) % P
            assert got[e] == acc
    else:
        # size-independent property: evaluating at a point of the LDE coset reproduces the LDE value
        w = pow(3, (P - 1) >> (log_n + 1), P)
        idx = [5, 12345, (1 << log_n) + 7, (2 << log_n) - 1]
        zs = [3 * pow(w, i, P) % P for i in idx]
        got = oracle.from_mont(poly_eval(coeffs, [0] * 4, oracle.to_mont(zs)))
        want = oracle.from_mont(lde.numpy()[0, idx])
        assert got == want
