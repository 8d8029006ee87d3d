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

    and regrouped BY COLUMN so that almost every multiplication has a constant operand and lands in an
    unreduced dot product (program.py DOT; one Montgomery reduction per column instead of one per term):

        sum_t a_t (T_c(x) - y_t) / (x - z g^off_t)
            = sum_c T_c(x) * [ sum_{t in c} a_t g^-off_t u[i - b off_t] ]  -  sum_off K_off u[i - b off],
        K_off = g^-off * sum_{t at off} a_t y_t.

    trace_terms: (column, offset, claimed value y, coefficient);  comp_terms: (column, y, coefficient).
    `Trace(u_col, -off)` is a row offset in TRACE units, i.e. -off * blowup LDE rows, exactly the shift above."""
    per_col: dict[int, Expr] = {}
    k_off: dict[int, int] = {}
    for col, off, y, coeff in trace_terms:
        gi = pow(g, -off, p)
        term = Constant(coeff * gi % p) * Trace(u_col, -off)
        per_col[col] = term if col not in per_col else per_col[col] + term
        k_off[off] = (k_off.get(off, 0) + coeff * y % p * gi) % p
    total = None
    for col, a in per_col.items():
        q = Trace(col, 0) * a
        total = q if total is None else total + q
    for off, k in k_off.items():
        if k:
            q = Constant(k) * Trace(u_col, -off)
            total = -q if total is None else total - q
    comp, comp_y = None, 0
    for col, y, coeff in comp_terms:
        term = Constant(coeff) * Trace(col, 0)
        comp = term if comp is None else comp + term
        comp_y = (comp_y + coeff * y) % p
    if comp is not None:
        q = (comp - Constant(comp_y)) * Trace(v_col, 0)
        total = q if total is None else total + q
    return total


def deep_terms(taps, ood_trace, ood_composition, first_comp_col: int, alpha: int, p: int):
    """Stark::gen_deep_coeffs (reference src/lib.rs:102-116): powers of one alpha, first over the trace arguments
    (the taps, in air.trace_arguments() order) then over the composition columns.  Returns (trace_terms, comp_terms)
    for deep_expr_shifted.  Note alpha^0 = 1: the first term has a unit coefficient, which is part of the program's
    structure (tools/gen_ce_kernels.py builds its terms with this function for that reason)."""
    t_terms, c_terms, k = [], [], 0
    for (col, off), y in zip(taps, ood_trace):
        t_terms.append((col, off, y, pow(alpha, k, p)))
        k += 1
    for j, y in enumerate(ood_composition):
        c_terms.append((first_comp_col + j, y, pow(alpha, k, p)))
        k += 1
    return t_terms, c_terms


def deep_expr_symbolic(taps, n_comp: int, first_comp_col: int, u_col: int, v_col: int, g: int, p: int) -> Expr:
    """deep_expr_shifted with the per-proof values left open, for `compile_template`: Challenge(0) = the DEEP alpha,
    Hint(k) = the k-th out-of-domain value (trace arguments in air.trace_arguments() order, then the composition columns).
    Same regrouping, so the patched program has the structure of the one compiled from values (alpha^0 folds to the literal 1
    in both)."""
    from .expr import Challenge, Hint

    alpha = Challenge(0)
    per_col: dict[int, Expr] = {}
    k_off: dict[int, Expr] = {}
    k = 0
    for col, off in taps:
        gi = Constant(pow(g, -off, p))
        coeff = alpha.pow(k) * gi
        term = coeff * Trace(u_col, -off)
        per_col[col] = term if col not in per_col else per_col[col] + term
        ky = coeff * Hint(k)
        k_off[off] = ky if off not in k_off else k_off[off] + ky
        k += 1
    total = None
    for col, a in per_col.items():
        q = Trace(col, 0) * a
        total = q if total is None else total + q
    for off, kk in k_off.items():
        q = kk * Trace(u_col, -off)
        total = -q if total is None else total - q
    comp, comp_y = None, None
    for j in range(n_comp):
        coeff = alpha.pow(k)
        term = coeff * Trace(first_comp_col + j, 0)
        comp = term if comp is None else comp + term
        cy = coeff * Hint(k)
        comp_y = cy if comp_y is None else comp_y + cy
        k += 1
    if comp is not None:
        q = (comp - comp_y) * Trace(v_col, 0)
        total = q if total is None else total + q
    return total


def deep_expr_filtered(taps, n_comp: int, first_comp_col: int, u_col: int, v_col: int, g: int, p: int, filter_cols: dict, value_col: int) -> Expr:
    """deep_expr_symbolic with the long sums taken out of the per-row program.  In the by-column form

        deep(x) = sum_c T_c(x) W_c(x) - V(x) + (composition part),
        W_c(x) = sum_{t in c} a_t / (x - z g^off_t),      V(x) = sum_t a_t y_t / (x - z g^off_t),

    W_c and V are sums of simple poles on z<g>.  On the evaluation coset every x has the same x^n = K, so
    1 / (x - zeta) = sum_k x^k zeta^(n-1-k) / (K - z^n) there: each of them is a POLYNOMIAL on the coset, whose n values cost two
    size-n transforms whatever the number of poles (prover.py `_pole_sum_on_coset`).  filter_cols: {trace column: working column
    that holds W_c on the coset rows} for the columns with many taps; value_col: the working column that holds V.  The columns
    with few taps keep their shifted reads of u.  Same polynomial, hence the same values as deep_expr_symbolic."""
    from .expr import Challenge

    alpha = Challenge(0)
    per_col: dict[int, Expr] = {}
    k = 0
    for col, off in taps:
        if col not in filter_cols:
            term = alpha.pow(k) * Constant(pow(g, -off, p)) * Trace(u_col, -off)
            per_col[col] = term if col not in per_col else per_col[col] + term
        k += 1
    total = None
    for col, a in per_col.items():
        q = Trace(col, 0) * a
        total = q if total is None else total + q
    for col in sorted(filter_cols):
        q = Trace(col, 0) * Trace(filter_cols[col], 0)
        total = q if total is None else total + q
    total = total - Trace(value_col, 0)
    from .expr import Hint

    comp, comp_y = None, None
    for j in range(n_comp):
        coeff = alpha.pow(k)
        term = coeff * Trace(first_comp_col + j, 0)
        comp = term if comp is None else comp + term
        cy = coeff * Hint(k)
        comp_y = cy if comp_y is None else comp_y + cy
        k += 1
    if comp is not None:
        total = total + (comp - comp_y) * Trace(v_col, 0)
    return total


DEEP_FILTER_MIN_TAPS = 40      # one transform pair costs about as much as 37 taps of the per-row form (prover.ProofOptions default)


def deep_filter_columns(taps, min_taps: int) -> list:
    """the trace columns whose W_c goes through a transform (at least min_taps mask offsets), in ascending order"""
    count: dict[int, int] = {}
    for col, _ in taps:
        count[col] = count.get(col, 0) + 1
    return sorted(col for col, c in count.items() if min_taps and c >= min_taps)
