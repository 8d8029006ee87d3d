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
