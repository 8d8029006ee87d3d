"""DEEP composition polynomial (ministark `DeepPolyComposer`, coefficients from
`Stark::gen_deep_coeffs`, reference src/lib.rs:102-116: powers of one alpha, first over the trace
arguments then over the composition columns, degree adjustment (1, 0)):

    deep(x) = sum_t  alpha_t * (T_{col_t}(x) - y_t) / (x - z_t)

where z_t = z * g^offset_t for trace terms and z^ce for composition-column terms, and y_t the claimed
out-of-domain value.  It is an ordinary expression over the LDE row (taps at offset 0 only, one
full-period denominator per distinct point), so it is compiled and executed by the same machinery as the
constraint composition: `compile_program(deep_expr(terms), ...)` + `evaluate(...)`; the distinct
denominators are inverted with one batched inversion per row."""
from __future__ import annotations

from .expr import Constant, Expr, Trace, X


def deep_expr(terms) -> Expr:
    """terms: iterable of (column, point z_t, claimed value y_t, coefficient alpha_t), canonical ints."""
    by_point: dict[int, Expr] = {}
    for col, z, y, coeff in terms:
        term = Constant(coeff) * (Trace(col, 0) - Constant(y))
        by_point[z] = term if z not in by_point else by_point[z] + term
    total = None
    for z, num in by_point.items():
        q = num / (X - Constant(z))
        total = q if total is None else total + q
    return total


def deep_expr_shifted(trace_terms, comp_terms, u_col: int, v_col: int, g: int, p: int) -> Expr:
    """Same polynomial, with every denominator read from two precomputed columns instead of inverted per row:

        u[i] = 1 / (x_i - z)      (column u_col)        v[i] = 1 / (x_i - z^ce)     (column v_col)
        1 / (x_i - z g^off) = g^-off * u[i - blowup * off]            (x_i - z g^off = g^off (x_{i - b off} - z))

    trace_terms: (column, offset, claimed value y, coefficient);  comp_terms: (column, y, coefficient).
    `Trace(u_col, -off)` is a row offset in TRACE units, i.e. -off * blowup LDE rows, exactly the shift above."""
    by_off: dict[int, Expr] = {}
    for col, off, y, coeff in trace_terms:
        term = Constant(coeff) * (Trace(col, 0) - Constant(y))
        by_off[off] = term if off not in by_off else by_off[off] + term
    total = None
    for off, num in by_off.items():
        q = num * Constant(pow(g, -off, p)) * Trace(u_col, -off)
        total = q if total is None else total + q
    comp = None
    for col, y, coeff in comp_terms:
        term = Constant(coeff) * (Trace(col, 0) - Constant(y))
        comp = term if comp is None else comp + term
    if comp is not None:
        q = comp * Trace(v_col, 0)
        total = q if total is None else total + q
    return total
