"""CPU-only: the Expr -> program compiler (period analysis, table extraction, slot allocation).
The program is executed by a tiny Python interpreter of the blob and compared with the independent
tree evaluator — this pins the blob format the CUDA kernel consumes."""
import numpy as np
import pytest

from air_ref import eval_expr
from sandstorm_b200.air import Challenge, Constant, Hint, Periodic, Trace, X, compile_program, composition_constraint
from sandstorm_b200.air.expr import P

from blob_emu import run_blob


def toy_air(n):
    """A small AIR exercising every leaf and operator kind the Cairo layouts use."""
    g = pow(3, (P - 1) // n, P)
    one = Constant(1)
    every_row_inv = one / (X.pow(n) - one)
    every4_inv = one / (X.pow(n // 4) - one)
    last_row = X - Constant(pow(g, n - 1, P))
    first_row_inv = one / (X - one)
    per = Periodic([5, 7, 11, 13], 8)
    c0 = (Trace(0, 0) * Trace(0, 0) - Trace(0, 1)) * every_row_inv * last_row
    c1 = (Trace(1, 1) - Trace(1, 0) * (Challenge(0) - Trace(0, 0) - Challenge(1) * Trace(2, 2))) * every4_inv
    c2 = (Trace(2, 0) - Hint(0)) * first_row_inv
    c3 = (Trace(1, 3) * per - Trace(0, 2).pow(3)) * every4_inv
    c4 = (Trace(0, 5) + Trace(1, 0) * Constant(2).pow(64) - Hint(1)) / last_row
    c5 = -(Trace(2, 1) - X * Trace(2, 0)) * (X.pow(n // 2) - Constant(pow(g, n // 2, P))) * every_row_inv
    return [c0, c1, c2, c3, c4, c5]


@pytest.mark.parametrize("log_n,log_blowup", [(3, 1), (4, 2), (6, 1)])
def test_compiled_program_matches_tree_evaluator(log_n, log_blowup):
    rng = np.random.default_rng(log_n)
    n, N = 1 << log_n, 1 << (log_n + log_blowup)
    lde_int = [[int.from_bytes(rng.bytes(31), "big") for _ in range(N)] for _ in range(3)]
    challenges = [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
    hints = [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
    alpha = [int.from_bytes(rng.bytes(31), "big")]
    expr = composition_constraint(toy_air(n))
    prog = compile_program(expr, log_n, log_blowup, challenges, hints, alpha)
    assert prog.n_tables >= 2 and prog.n_batch_inv == 2          # X - 1 and X - g^(n-1)
    assert prog.n_slots <= 32
    for i in list(range(min(N, 16))) + [N - 1, N // 2 + 1]:
        want = eval_expr(expr, i, lde_int, log_n, log_blowup, challenges, hints, alpha)
        assert run_blob(prog.blob, i, lde_int, log_n + log_blowup) == want, i


def test_period_classification():
    log_n, log_b = 5, 1
    n = 1 << log_n
    e = (Trace(0, 0) - Constant(3)) / (X.pow(n // 4) - Constant(1)) + Periodic([1, 2], 2) * X.pow(n)
    prog = compile_program(e, log_n, log_b)
    # 1/(X^(n/4) - 1): period 8; Periodic(.,2) * X^n: periods 4 and 2 -> one table of 4
    assert sorted(prog.table_sizes) == [4, 8]
    assert prog.n_batch_inv == 0 and prog.n_trace_taps == 1


def test_constant_folding_and_cse():
    e = (Trace(0, 0) + Constant(2) * Constant(3)) * (Trace(0, 0) + Constant(6)) + Challenge(0).pow(5)
    prog = compile_program(e, 3, 1, challenges=[2])
    # (t + 6) is shared, Challenge^5 folds to the constant 32
    assert prog.n_trace_taps == 1 and prog.n_mul == 1


def test_small_coefficients_become_addition_chains():
    """|c| <= 16 costs raw additions, not a multiplication; the emulator checks every bound on the way."""
    rng = np.random.default_rng(8)
    log_n, log_b = 4, 1
    N = 1 << (log_n + log_b)
    lde_int = [[int.from_bytes(rng.bytes(31), "big") for _ in range(N)] for _ in range(3)]
    a, b, c = Trace(0, 0), Trace(1, 1), Trace(2, 0)
    e = (Constant(2) * a - Constant(3) * b + Constant(16) * c * a - Constant(10) * (a * b) + Constant(P - 6) * c) * (a - Constant(13) * b)
    prog = compile_program(e, log_n, log_b)
    assert prog.n_mul == 3                                   # c*a, a*b and the outer product only
    for i in range(N):
        assert run_blob(prog.blob, i, lde_int, log_n + log_b) == eval_expr(e, i, lde_int, log_n, log_b, [], [], [0])


def test_structure_hash_ignores_size_and_values():
    """The hash that selects a build-time specialised kernel depends on the program's structure only: not on the
    trace length, the challenge draw or the constants (tools/gen_ce_kernels.py relies on this)."""
    import random

    from sandstorm_b200.air.layouts import load_layout
    from sandstorm_b200.air.program import structure_hash, tap_reach

    L = load_layout("recursive")
    hashes = set()
    for log_n, seed in ((13, 1), (14, 2), (16, 3)):
        rnd = random.Random(seed)
        prog = compile_program(L.composition(1 << log_n, inv_x_minus_one_col=L.num_columns + 2), log_n, 1, [rnd.randrange(P) for _ in range(L.n_challenges())],
                               [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)], with_tables=False)
        hashes.add(structure_hash(prog.blob))
        reach = tap_reach(prog.blob, log_n + 1)
        assert max(hi for _, hi in reach.values()) == 2 * L.max_offset and min(lo for lo, _ in reach.values()) >= -2 * 16
    assert len(hashes) == 1
    other = compile_program(L.composition(1 << 13), 13, 1, [1] * L.n_challenges(), [2] * L.n_hints(), [3], with_tables=False)
    assert structure_hash(other.blob) not in hashes


def test_deep_quotient_program_matches_definition():
    """deep_expr_shifted (regrouped by column, shifted reads of u = 1/(x - z), v = 1/(x - z^ce)) == the definition
    sum a_t (T(x) - y_t) / (x - z g^off), row by row, through the emulator of the device interpreter."""
    import random

    from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms

    rnd = random.Random(3)
    log_n, b = 5, 1
    n, N = 1 << log_n, 1 << (log_n + b)
    g, w = pow(3, (P - 1) // n, P), pow(3, (P - 1) // N, P)
    ncol, ce = 3, 2
    lde = [[rnd.randrange(P) for _ in range(N)] for _ in range(ncol + ce)]
    z, alpha = rnd.randrange(P), rnd.randrange(P)
    zc = pow(z, ce, P)
    taps = [(c, off) for c in range(ncol) for off in (0, 1, 2, 5, 7)[: 3 + c]]
    tt, ct = deep_terms(taps, [rnd.randrange(P) for _ in taps], [rnd.randrange(P) for _ in range(ce)], ncol, alpha, P)
    u = [pow(3 * pow(w, i, P) - z, -1, P) for i in range(N)]
    v = [pow(3 * pow(w, i, P) - zc, -1, P) for i in range(N)]
    prog = compile_program(deep_expr_shifted(tt, ct, ncol + ce, ncol + ce + 1, g, P), log_n, b)
    assert prog.n_dot >= ncol
    for i in range(N):
        x = 3 * pow(w, i, P) % P
        want = (sum(a * (lde[c][i] - y) * pow(x - z * pow(g, off, P), -1, P) for c, off, y, a in tt)
                + sum(a * (lde[c][i] - y) * pow(x - zc, -1, P) for c, y, a in ct)) % P
        assert run_blob(prog.blob, i, lde + [u, v], log_n + b) == want


@pytest.mark.parametrize("name,log_n", [("plain", 8), ("recursive", 15), ("starknet", 19)])
def test_generated_kernel_registry_matches_prover_programs(name, log_n):
    """The registry written by tools/gen_ce_kernels.py (csrc/ce_gen.cuh, built by __graft_entry__.build()) must contain the
    structure hash of the programs the prover compiles — at another trace length and challenge draw than the generator's —
    otherwise ss_constraint_eval silently runs the interpreter instead of the specialised kernels."""
    import os
    import random
    import re

    from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms
    from sandstorm_b200.air.layouts import load_layout
    from sandstorm_b200.air.program import structure_hash

    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sandstorm_b200", "csrc", "ce_gen.cuh")
    if not os.path.exists(path):
        import __graft_entry__ as g

        g.build()
    registry = {int(h, 16): n for h, n in re.findall(r"\{0x([0-9a-f]{16})ull, ce_gen_(\w+),", open(path).read())}
    L = load_layout(name)
    C, ce, n = L.num_columns, 2, 1 << log_n
    rnd = random.Random(991 + log_n)
    from sandstorm_b200.air import compile_template

    # (the prover compiles the composition as a template and patches the challenges in; hints of a real proof are 0 / 1 / small)
    comp = compile_template(L.composition(n, inv_x_minus_one_col=C + ce), log_n, 1, L.n_challenges(), L.n_hints(), 1, with_tables=False)
    comp = comp.patch([rnd.randrange(P) for _ in range(L.n_challenges())], [rnd.randrange(3) for _ in range(L.n_hints())], [rnd.randrange(P)])
    tt, ct = deep_terms(L.taps(), [rnd.randrange(P) for _ in L.taps()], [rnd.randrange(P) for _ in range(ce)], C, rnd.randrange(P), P)
    deep = compile_program(deep_expr_shifted(tt, ct, C + ce + 1, C + ce + 2, pow(3, (P - 1) // n, P), P), log_n, 1, with_tables=False)
    assert registry.get(structure_hash(comp.blob), "").startswith(f"{name}_composition")
    assert registry.get(structure_hash(deep.blob), "").startswith(f"{name}_deep")
    # the form the prover runs when the layout has long pole sums (prover.HotPathProver.__init__ decides the same way)
    from sandstorm_b200.air.deep import DEEP_FILTER_MIN_TAPS, deep_expr_filtered, deep_filter_columns

    taps = L.taps()
    heavy = deep_filter_columns(taps, DEEP_FILTER_MIN_TAPS)
    if heavy or len({off for _, off in taps}) >= DEEP_FILTER_MIN_TAPS:
        value_col = C + ce + 3
        fcols = {col: value_col + 1 + j for j, col in enumerate(heavy)}
        t = compile_template(deep_expr_filtered(taps, ce, C, C + ce + 1, C + ce + 2, pow(3, (P - 1) // n, P), P, fcols, value_col), log_n, 1, 1,
                             len(taps) + ce, 1, with_tables=False)
        deepf = t.patch([rnd.randrange(P)], [rnd.randrange(P) for _ in range(len(taps) + ce)], [0])
        assert registry.get(structure_hash(deepf.blob), "").startswith(f"{name}_deepf")
    else:
        assert name == "plain"


@pytest.mark.parametrize("special_hints", [False, True])
def test_template_patch_equals_direct_evaluation(special_hints):
    """compile_template (tracked challenges, air/symbolic.py) + patch(values): the patched program evaluates the
    expression for those values — for generic draws and for hints that are 0 / 1 (which a direct compilation would
    fold into a different structure) — and its structure does not depend on the values."""
    from sandstorm_b200.air import compile_template
    from sandstorm_b200.air.program import structure_hash

    rng = np.random.default_rng(77)
    log_n, log_blowup = 4, 1
    n, N = 1 << log_n, 1 << (log_n + log_blowup)
    lde_int = [[int.from_bytes(rng.bytes(31), "big") for _ in range(N)] for _ in range(3)]
    expr = composition_constraint(toy_air(n))
    tpl = compile_template(expr, log_n, log_blowup, n_challenges=2, n_hints=2, n_coeffs=1)
    hashes = set()
    for draw in range(2):
        challenges = [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
        hints = [draw, 1 - draw] if special_hints else [int.from_bytes(rng.bytes(31), "big") for _ in range(2)]
        alpha = [int.from_bytes(rng.bytes(31), "big")]
        prog = tpl.patch(challenges, hints, alpha)
        hashes.add(structure_hash(prog.blob))
        for i in list(range(8)) + [N - 1, N // 2 + 1]:
            assert run_blob(prog.blob, i, lde_int, log_n + log_blowup) == eval_expr(expr, i, lde_int, log_n, log_blowup, challenges, hints, alpha), i
    assert len(hashes) == 1
    with pytest.raises(ValueError):
        tpl.patch([1], [2, 3], [4])


def test_template_of_a_real_layout_matches_tree_evaluator():
    """recursive layout at its minimum trace length: template + patch through the device emulator == the big-int tree
    evaluator of the same constraints (challenge-dependent table factors are multiplied out the way the device does)."""
    import random

    from sandstorm_b200.air import compile_template
    from sandstorm_b200.air.layouts import load_layout

    L = load_layout("recursive")
    log_n, b = 11, 1
    n, N = 1 << log_n, 1 << (log_n + b)
    rnd = random.Random(5)
    C = L.num_columns
    lde = [[rnd.randrange(P) for _ in range(N)] for _ in range(C)]
    w = pow(3, (P - 1) // N, P)
    w_col = [pow(3 * pow(w, i, P) - 1, -1, P) for i in range(N)]
    tpl = compile_template(L.composition(n, inv_x_minus_one_col=C), log_n, b, L.n_challenges(), L.n_hints(), 1)
    ch, hints, alpha = [rnd.randrange(P) for _ in range(L.n_challenges())], [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)]
    hints[5], hints[8], hints[9] = 1, 1, 0                       # RangeCheckProduct, DilutedCheckProduct, DilutedCheckFirst of a real proof
    prog = tpl.patch(ch, hints, alpha)
    expr = L.composition(n)
    for i in (0, 1, 5, 2047, N - 3):
        assert run_blob(prog.blob, i, lde + [w_col], log_n + b) == eval_expr(expr, i, lde, log_n, b, ch, hints, alpha), i


@pytest.mark.parametrize("name,log_n", [("plain", 7), ("recursive", 12), ("starknet", 16)])
def test_deep_template_patch_is_the_direct_compilation(name, log_n):
    """the DEEP quotient program compiled once with alpha and the out-of-domain values open, then patched, is byte for byte
    the program compiled from the values (whose semantics test_deep_quotient_program_matches_definition pins)."""
    import random

    from sandstorm_b200.air import compile_template
    from sandstorm_b200.air.deep import deep_expr_shifted, deep_expr_symbolic, deep_terms
    from sandstorm_b200.air.layouts import load_layout

    L = load_layout(name)
    n, C, ce = 1 << log_n, L.num_columns, 2
    g, taps, rnd = pow(3, (P - 1) // n, P), L.taps(), random.Random(log_n)
    tpl = compile_template(deep_expr_symbolic(taps, ce, C, C + ce + 1, C + ce + 2, g, P), log_n, 1, 1, len(taps) + ce, 1)
    for _ in range(2):
        ood, oc, alpha = [rnd.randrange(P) for _ in taps], [rnd.randrange(P) for _ in range(ce)], rnd.randrange(P)
        tt, ct = deep_terms(taps, ood, oc, C, alpha, P)
        assert tpl.patch([alpha], ood + oc, [0]).blob == compile_program(deep_expr_shifted(tt, ct, C + ce + 1, C + ce + 2, g, P), log_n, 1).blob


def _pole_sum_by_transforms(weights, z, n, g):
    """prover._pole_sums_on_coset in big ints: inverse transform of the sparse vector weights[off] g^-off, scaling by
    C z^(n-1) (3/z)^k, forward transform.  -> [sum_off weights[off] / (3 g^i - z g^off) for i < n]."""
    K, zn = pow(3, n, P), pow(z, n, P)
    c0 = pow((K - zn) % P, -1, P) * pow(z, n - 1, P) % P
    h0 = 3 * pow(z, -1, P) % P
    s = [0] * n
    for off, wgt in weights.items():
        s[off % n] = (s[off % n] + wgt * pow(g, -off, P)) % P
    ginv = pow(g, -1, P)
    B = [sum(s[o] * pow(ginv, o * k, P) for o in range(n) if s[o]) % P for k in range(n)]
    D = [c0 * pow(h0, k, P) % P * B[k] % P for k in range(n)]
    return [sum(D[k] * pow(g, i * k, P) for k in range(n)) % P for i in range(n)]


def test_pole_sums_are_polynomials_on_the_coset():
    """the identity behind the transform form of the DEEP quotient: on 3<g>, sum_off w_off / (x - z g^off) equals a polynomial whose
    coefficients are C z^(n-1) (3/z)^k B[k] — checked against the definition for several pole sets (incl. offsets >= n, duplicates)."""
    import random

    rnd = random.Random(21)
    for log_n in (3, 5):
        n = 1 << log_n
        g = pow(3, (P - 1) // n, P)
        z = rnd.randrange(P)
        for offs in ([0], [1, 5], [0, 2, 3, n + 1, 7], list(range(n))):
            weights = {o: rnd.randrange(P) for o in offs}
            got = _pole_sum_by_transforms(weights, z, n, g)
            for i in range(n):
                x = 3 * pow(g, i, P) % P
                want = sum(wgt * pow((x - z * pow(g, o, P)) % P, -1, P) for o, wgt in weights.items()) % P
                assert got[i] == want, (log_n, offs, i)


def test_filtered_deep_program_matches_definition():
    """deep_expr_filtered (V and the tap-heavy columns' W_c read from precomputed columns holding the pole sums, the other
    columns' poles as shifted reads of u) == the definition of the DEEP quotient on the sub-coset rows, through the emulator of the
    device interpreter; the pole-sum columns are filled by the transform identity above."""
    import random

    from sandstorm_b200.air import compile_template
    from sandstorm_b200.air.deep import deep_expr_filtered, deep_filter_columns

    rnd = random.Random(4)
    log_n, b = 4, 1
    n, N = 1 << log_n, 1 << (log_n + b)
    g, w = pow(3, (P - 1) // n, P), pow(3, (P - 1) // N, P)
    ncol, ce = 3, 2
    lde = [[rnd.randrange(P) for _ in range(N)] for _ in range(ncol + ce)]
    z, alpha = rnd.randrange(P), rnd.randrange(P)
    zc = pow(z, ce, P)
    taps = [(0, 0), (0, 1), (1, 0), (1, 1), (1, 2), (1, 5), (1, 7), (2, 3)]                  # column 1 is "heavy" at min_taps = 4
    assert deep_filter_columns(taps, 4) == [1]
    ood = [rnd.randrange(P) for _ in taps] + [rnd.randrange(P) for _ in range(ce)]
    u_col, v_col, value_col, f_col = ncol + ce, ncol + ce + 1, ncol + ce + 2, ncol + ce + 3
    tpl = compile_template(deep_expr_filtered(taps, ce, ncol, u_col, v_col, g, P, {1: f_col}, value_col), log_n, b, 1, len(taps) + ce, 1)
    prog = tpl.patch([alpha], ood, [0])
    u = [pow(3 * pow(w, i, P) - z, -1, P) for i in range(N)]
    v = [pow(3 * pow(w, i, P) - zc, -1, P) for i in range(N)]
    v_w, c_w, a = {}, {}, 1
    for (col, off), y in zip(taps, ood):
        v_w[off] = (v_w.get(off, 0) + a * y) % P
        if col == 1:
            c_w[off] = (c_w.get(off, 0) + a) % P
        a = a * alpha % P
    V, W1 = [0] * N, [0] * N
    V[::2] = _pole_sum_by_transforms(v_w, z, n, g)                                            # rows i << b of the working matrix
    W1[::2] = _pole_sum_by_transforms(c_w, z, n, g)
    for i in range(0, N, 2):
        x = 3 * pow(w, i, P) % P
        want, a = 0, 1
        for (col, off), y in zip(taps, ood):
            want = (want + a * (lde[col][i] - y) * pow(x - z * pow(g, off, P), -1, P)) % P
            a = a * alpha % P
        for j in range(ce):
            want = (want + a * (lde[ncol + j][i] - ood[len(taps) + j]) * pow(x - zc, -1, P)) % P
            a = a * alpha % P
        assert run_blob(prog.blob, i, lde + [u, v, V, W1], log_n + b) == want
