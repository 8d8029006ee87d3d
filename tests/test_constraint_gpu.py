"""GPU parity: ss_constraint_eval vs the independent big-int tree evaluator on the LDE of a random trace."""
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


@pytest.mark.parametrize("log_n,log_blowup", [(3, 1), (5, 2), (10, 1), (13, 1)])
def test_toy_air_composition_matches_tree_evaluator(ss, oracle, log_n, log_blowup):
    import torch

    from air_ref import eval_expr
    from sandstorm_b200.air import compile_program, composition_constraint
    from sandstorm_b200.air.evaluate import evaluate
    from test_air_compile import toy_air

    rng = np.random.default_rng(900 + log_n)
    n, N = 1 << log_n, 1 << (log_n + log_blowup)
    trace = oracle.random_felts(rng, 3, n)
    lde = ss.Matrix.from_numpy(trace).lde(log_blowup)
    lde_np = lde.numpy()
    assert np.array_equal(lde_np, oracle.lde(trace, log_blowup))
    challenges = [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
    hints = [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
    alpha = [int.from_bytes(rng.bytes(31), "big")]
    expr = composition_constraint(toy_air(n))
    prog = compile_program(expr, log_n, log_blowup, challenges, hints, alpha)
    got = evaluate(prog, lde, log_blowup)
    torch.cuda.synchronize()
    got_int = oracle.from_mont(got.cpu().numpy().view(np.uint64))
    rows = list(range(N)) if N <= 128 else [0, 1, 2, 3, 5, 8, N // 2, N // 2 + 1, N - 9, N - 2, N - 1] + [int(x) for x in rng.integers(0, N, 12)]
    cols_int = None
    if N <= 128:
        cols_int = [oracle.from_mont(lde_np[c]) for c in range(3)]
    for i in rows:
        if cols_int is None:
            # only the tapped rows are needed: offsets 0..5 (x blowup)
            need = sorted({(i + k * (1 << log_blowup)) % N for k in range(6)})
            sparse = [dict(zip(need, oracle.from_mont(lde_np[c][need]))) for c in range(3)]

            class Col(dict):
                def __getitem__(self, k):
                    return dict.__getitem__(self, k % N)

            view = [Col(s) for s in sparse]
        else:
            view = cols_int
        assert got_int[i] == eval_expr(expr, i, view, log_n, log_blowup, challenges, hints, alpha), i


def test_rejects_malformed_programs(ss, oracle):
    import ctypes

    import torch

    from sandstorm_b200.air import Trace, compile_program

    prog = compile_program(Trace(0, 0) * Trace(0, 1), 3, 1)
    d = torch.zeros((1, 16, 4), dtype=torch.int64, device="cuda")
    out = torch.zeros((16, 4), dtype=torch.int64, device="cuda")
    c = ss.default_context()
    bad = bytearray(prog.blob)
    bad[0] ^= 0xFF
    assert c.lib.ss_constraint_eval(c.handle, bytes(bad), len(bad), ctypes.c_void_p(d.data_ptr()), 16, 1, 3, 1, 0, 0, 0, ctypes.c_void_p(out.data_ptr()), None) == -1
    # wrong size / wrong log_n
    assert c.lib.ss_constraint_eval(c.handle, prog.blob, len(prog.blob) - 32, ctypes.c_void_p(d.data_ptr()), 16, 1, 3, 1, 0, 0, 0, ctypes.c_void_p(out.data_ptr()), None) == -1
    assert c.lib.ss_constraint_eval(c.handle, prog.blob, len(prog.blob), ctypes.c_void_p(d.data_ptr()), 32, 1, 4, 1, 0, 0, 0, ctypes.c_void_p(out.data_ptr()), None) == -1


def test_deep_composition_matches_definition(ss, oracle):
    """DEEP quotient over the LDE domain: with y_t the true evaluations the result is a polynomial of
    degree < n - 1 (checked through the inverse NTT), and every sampled row matches the big-int formula."""
    import torch

    from sandstorm_b200.air import compile_program
    from sandstorm_b200.air.deep import deep_expr
    from sandstorm_b200.air.evaluate import evaluate
    from sandstorm_b200.matrix import poly_eval

    P = oracle.P
    rng = np.random.default_rng(77)
    log_n, log_b, n_cols = 8, 1, 4
    n, N = 1 << log_n, 1 << (log_n + log_b)
    trace = oracle.random_felts(rng, n_cols, n)
    lde, coeffs = ss.Matrix.from_numpy(trace).lde(log_b, keep_coeffs=True)
    g = pow(3, (P - 1) // n, P)
    z = int.from_bytes(rng.bytes(31), "big")
    alpha = int.from_bytes(rng.bytes(31), "big")
    offsets = [0, 1, 2, 5, 16, 17, 33, 100, 255] + list(range(40, 90))          # > 32 distinct points: exercises the batch limit
    taps = [(c, off) for c in range(n_cols) for off in offsets[: 20 + 10 * c]]
    pts = [z * pow(g, off, P) % P for _, off in taps]
    ys = oracle.from_mont(poly_eval(coeffs, [c for c, _ in taps], oracle.to_mont(pts)))
    terms = [(c, pt, y, pow(alpha, k, P)) for k, ((c, _), pt, y) in enumerate(zip(taps, pts, ys))]
    prog = compile_program(deep_expr(terms), log_n, log_b)
    assert prog.n_batch_inv == len(set(pts))
    got = evaluate(prog, lde, log_b)
    torch.cuda.synchronize()
    got_np = got.cpu().numpy().view(np.uint64)
    lde_int = [oracle.from_mont(col) for col in lde.numpy()]
    w = pow(3, (P - 1) // N, P)
    for i in [0, 1, 7, N // 2, N - 1, 123]:
        x = 3 * pow(w, i, P) % P
        want = sum(a * (lde_int[c][i] - y) * pow(x - pt, -1, P) for c, pt, y, a in terms) % P
        assert oracle.from_mont(got_np[i:i + 1])[0] == want
    # low-degree check: interpolate the N evaluations on the coset; coefficients >= n - 1 must vanish
    m = ss.Matrix(got.reshape(1, N, 4).contiguous())
    c = oracle.from_mont(m.ntt_(inverse=True, coset=True).numpy()[0])
    assert all(v == 0 for v in c[n - 1:]) and any(v != 0 for v in c[: n - 1])


@pytest.mark.parametrize("name,log_n", [("recursive", 12), ("starknet", 15)])
def test_layout_composition_on_gpu(ss, oracle, name, log_n):
    """The real Cairo AIRs (transpiled from layouts/src/*/air.rs): GPU evaluation of the composition
    constraint on the LDE of a random trace vs the independent tree evaluator at sampled rows."""
    import random

    import torch

    from air_ref import eval_expr
    from sandstorm_b200.air import compile_program
    from sandstorm_b200.air.evaluate import evaluate
    from sandstorm_b200.air.layouts import load_layout

    L = load_layout(name)
    rnd = random.Random(log_n)
    P = oracle.P
    n, N = 1 << log_n, 2 << log_n
    rng = np.random.default_rng(log_n)
    lde_np = oracle.random_felts(rng, L.num_columns, N)           # any matrix: the evaluator does not need a valid trace
    lde = ss.Matrix.from_numpy(lde_np)
    ch = [rnd.randrange(P) for _ in range(L.n_challenges())]
    hints = [rnd.randrange(P) for _ in range(L.n_hints())]
    alpha = [rnd.randrange(P)]
    expr = L.composition(n)
    prog = compile_program(expr, log_n, 1, ch, hints, alpha)
    got = evaluate(prog, lde, 1)
    torch.cuda.synchronize()
    got_np = got.cpu().numpy().view(np.uint64)
    taps = L.taps()
    for i in [0, 1, 2, 15, 16, 31, N // 2, N - 1, N - 2 * L.max_offset - 1] + [rnd.randrange(N) for _ in range(6)]:
        i %= N
        cols = [dict() for _ in range(L.num_columns)]
        for c, off in taps:
            r = (i + 2 * off) % N
            cols[c][r] = oracle.from_mont(lde_np[c][r:r + 1])[0]
        assert oracle.from_mont(got_np[i:i + 1])[0] == eval_expr(expr, i, cols, log_n, 1, ch, hints, alpha), (name, i)


def test_shifted_inverse_rewrite_is_equivalent(ss, oracle):
    """1/(x_i - g^e) = g^-e * w[i - 2e] with w = 1/(x - 1): the rewritten recursive-layout program (no per-row
    inversion) gives the same composition evaluations as the direct one; and ss_inv_x_minus_c matches big ints."""
    import random

    import torch

    from sandstorm_b200.air import compile_program
    from sandstorm_b200.air.evaluate import evaluate
    from sandstorm_b200.air.layouts import load_layout
    from sandstorm_b200.matrix import inv_x_minus_c

    P = oracle.P
    L = load_layout("recursive")
    log_n = 12
    n, N = 1 << log_n, 2 << log_n
    rnd = random.Random(4)
    rng = np.random.default_rng(4)
    cols = torch.zeros((L.num_columns + 1, N, 4), dtype=torch.int64, device="cuda")
    cols[: L.num_columns] = torch.from_numpy(oracle.random_felts(rng, L.num_columns, N).view(np.int64)).cuda()
    inv_x_minus_c(cols[L.num_columns], oracle.to_mont([1])[0])
    w = pow(3, (P - 1) // N, P)
    got_w = oracle.from_mont(cols[L.num_columns].cpu().numpy().view(np.uint64))
    for i in (0, 1, 5, N // 2, N - 1):
        assert got_w[i] == pow(3 * pow(w, i, P) - 1, -1, P)
    ch, hints, alpha = [rnd.randrange(P) for _ in range(6)], [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)]
    m = ss.Matrix(cols)
    direct = evaluate(compile_program(L.composition(n), log_n, 1, ch, hints, alpha), m, 1)
    prog = compile_program(L.composition(n, inv_x_minus_one_col=L.num_columns), log_n, 1, ch, hints, alpha)
    assert prog.n_batch_inv == 0
    shifted = evaluate(prog, m, 1)
    torch.cuda.synchronize()
    assert torch.equal(direct, shifted)


@pytest.mark.parametrize("name,log_n", [("recursive", 13), ("starknet", 17)])
def test_specialised_kernels_match_interpreter(ss, oracle, name, log_n):
    """The build-time specialisations (tools/gen_ce_kernels.py) of the layout's composition and DEEP programs are
    selected by structure hash — for any trace length and any challenge draw — and give bit-identical results to
    the interpreter, which is itself checked against the tree evaluator above."""
    import random

    import torch

    from sandstorm_b200.air import compile_program
    from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms
    from sandstorm_b200.air.evaluate import evaluate
    from sandstorm_b200.air.layouts import load_layout

    P = oracle.P
    L = load_layout(name)
    C, ce = L.num_columns, 2
    n, N = 1 << log_n, 2 << log_n
    rnd = random.Random(log_n)
    rng = np.random.default_rng(log_n)
    cols = torch.from_numpy(oracle.random_felts(rng, C + ce + 3, N).view(np.int64)).cuda()
    m = ss.Matrix(cols)
    c = m.ctx
    from sandstorm_b200.air import compile_template

    # the prover's path: challenge-independent template + value patch (hints of a real proof are small: 0, 1, addresses)
    comp = compile_template(L.composition(n, inv_x_minus_one_col=C + ce), log_n, 1, L.n_challenges(), L.n_hints(), 1).patch(
        [rnd.randrange(P) for _ in range(L.n_challenges())], [rnd.randrange(3) for _ in range(L.n_hints())], [rnd.randrange(P)])
    g = pow(3, (P - 1) // n, P)
    tt, ct = deep_terms(L.taps(), [rnd.randrange(P) for _ in L.taps()], [rnd.randrange(P) for _ in range(ce)], C, rnd.randrange(P), P)
    deep = compile_program(deep_expr_shifted(tt, ct, C + ce + 1, C + ce + 2, g, P), log_n, 1)
    try:
        for prog in (comp, deep):
            c.check(c.lib.ss_set_option(c.handle, b"ce_aot", 0))
            want = evaluate(prog, m, 1)
            assert c.lib.ss_get_option(c.handle, b"ce_last_aot", -1) == 0
            c.check(c.lib.ss_set_option(c.handle, b"ce_aot", 1))
            got = evaluate(prog, m, 1)
            assert c.lib.ss_get_option(c.handle, b"ce_last_aot", -1) == 1, "no specialised kernel matched the program's structure hash"
            torch.cuda.synchronize()
            assert torch.equal(want, got)
            # a row range only (the multi-GPU path) writes exactly that range
            part = torch.zeros_like(got)
            evaluate(prog, m, 1, out=part, rows=(N // 4, N // 2))
            torch.cuda.synchronize()
            assert torch.equal(part[N // 4:3 * N // 4], want[N // 4:3 * N // 4]) and not part[:N // 4].any() and not part[3 * N // 4:].any()
    finally:
        c.check(c.lib.ss_set_option(c.handle, b"ce_aot", 1))
