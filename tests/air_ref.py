"""Independent big-int evaluator of an Expr tree at one LDE row (test oracle for the constraint
evaluator; follows the definitions, shares no code with sandstorm_b200/air/program.py)."""
from sandstorm_b200.air.expr import P


def eval_expr(e, i, lde_int, log_n, log_blowup, challenges, hints, coeffs, memo=None):
    """lde_int[col][row]: canonical ints of the LDE matrix; x_i = 3 * w_N^i."""
    memo = {} if memo is None else memo
    n, N = 1 << log_n, 1 << (log_n + log_blowup)
    w = pow(3, (P - 1) // N, P)
    x = 3 * pow(w, i, P) % P

    def go(e):
        if e in memo:
            return memo[e]
        op = e.op
        if op == "x": v = x
        elif op == "const": v = e.args[0]
        elif op == "trace": v = lde_int[e.args[0]][(i + e.args[1] * (1 << log_blowup)) % N]
        elif op == "challenge": v = challenges[e.args[0]]
        elif op == "hint": v = hints[e.args[0]]
        elif op == "composition_coeff": v = coeffs[e.args[0]]
        elif op == "periodic":
            cs, interval = e.args
            y = pow(x, n // interval, P)
            v = sum(c * pow(y, k, P) for k, c in enumerate(cs)) % P
        elif op == "pow": v = pow(go(e.args[0]), e.args[1], P)
        elif op == "neg": v = -go(e.args[0]) % P
        else:
            a, b = go(e.args[0]), go(e.args[1])
            v = (a + b if op == "add" else a - b if op == "sub" else a * b if op == "mul" else a * pow(b, -1, P)) % P
        memo[e] = v
        return v

    return go(e)


_G = {}


def eval_fraction(e, i, trace_int, log_n, challenges, hints, memo=None, x=None):
    """The constraint `e` as a rational function (numerator, denominator) at row i of the TRACE domain (x = g^i, taps read
    trace_int[col][(i + off) mod n]).  A constraint holds iff its numerator vanishes wherever its denominator (the
    zerofier) does: no division is ever carried out, so the zerofier rows themselves can be evaluated."""
    memo = {} if memo is None else memo
    n = 1 << log_n
    if x is None:
        if n not in _G:
            _G[n] = pow(3, (P - 1) // n, P)
        x = pow(_G[n], i, P)

    def go(e):
        if e in memo:
            return memo[e]
        op = e.op
        if op == "x": v = (x, 1)
        elif op == "const": v = (e.args[0], 1)
        elif op == "trace": v = (trace_int[e.args[0]][(i + e.args[1]) % n], 1)
        elif op == "challenge": v = (challenges[e.args[0]], 1)
        elif op == "hint": v = (hints[e.args[0]], 1)
        elif op == "periodic":
            cs, interval = e.args
            y = pow(x, n // interval, P)
            v = (sum(c * pow(y, k, P) for k, c in enumerate(cs)) % P, 1)
        elif op == "pow":
            a, b = go(e.args[0])
            v = (pow(a, e.args[1], P), pow(b, e.args[1], P))
        elif op == "neg":
            a, b = go(e.args[0])
            v = (-a % P, b)
        else:
            (a, b), (c, d) = go(e.args[0]), go(e.args[1])
            if op == "add": v = ((a * d + c * b) % P, b * d % P) if b != d else ((a + c) % P, b)
            elif op == "sub": v = ((a * d - c * b) % P, b * d % P) if b != d else ((a - c) % P, b)
            elif op == "mul": v = (a * c % P, b * d % P)
            elif op == "div": v = (a * d % P, b * c % P)
            else: raise ValueError(op)
        memo[e] = v
        return v

    return go(e)


def divisors(e, out=None, seen=None):
    """every sub-expression something is divided by (the zerofiers): a constraint's denominator vanishes exactly where one
    of them does."""
    out = [] if out is None else out
    seen = set() if seen is None else seen
    if e in seen:
        return out
    seen.add(e)
    if e.op == "div" and e.args[1] not in out:
        out.append(e.args[1])
    for a in e.args:
        if hasattr(a, "op"):
            divisors(a, out, seen)
    return out


def eval_at_point(e, z, tap_values, log_n, challenges, hints, coeffs):
    """The verifier's evaluation of an Expr at an out-of-domain point z: Trace(col, off) reads the claimed value
    tap_values[(col, off)] = T_col(z * g^off) (ministark's `Verifier` recomputes the composition this way)."""
    n = 1 << log_n
    memo = {}

    def go(e):
        if e in memo:
            return memo[e]
        op = e.op
        if op == "x": v = z
        elif op == "const": v = e.args[0]
        elif op == "trace": v = tap_values[(e.args[0], e.args[1])]
        elif op == "challenge": v = challenges[e.args[0]]
        elif op == "hint": v = hints[e.args[0]]
        elif op == "composition_coeff": v = coeffs[e.args[0]]
        elif op == "periodic":
            cs, interval = e.args
            y = pow(z, n // interval, P)
            v = 0
            for c in reversed(cs):
                v = (v * y + c) % P
        elif op == "pow": v = pow(go(e.args[0]), e.args[1], P)
        elif op == "neg": v = -go(e.args[0]) % P
        else:
            a, b = go(e.args[0]), go(e.args[1])
            v = (a + b if op == "add" else a - b if op == "sub" else a * b if op == "mul" else a * pow(b, -1, P)) % P
        memo[e] = v
        return v

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    return go(e)
