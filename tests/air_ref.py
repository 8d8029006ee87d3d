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
            v = {"add": a + b, "sub": a - b, "mul": a * b, "div": a * pow(b, -1, P)}[op] % P
        memo[e] = v
        return v

    return go(e)
