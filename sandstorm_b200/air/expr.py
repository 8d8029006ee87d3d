"""Symbolic expressions over Fp252 — the mirror of ministark::expression::Expr with
AlgebraicItem leaves.  Nodes are immutable and hash-consed, which gives the effect of
`.reuse_shared_nodes()` (layouts/src/recursive/air.rs:1197) for free."""
from __future__ import annotations

P = 2**251 + 17 * 2**192 + 1


class Expr:
    __slots__ = ("op", "args", "_hash")
    _table: dict = {}

    def __new__(cls, op, *args):
        # (a tracked constant, air/symbolic.py, is keyed by its tape so that it never aliases a plain integer)
        key = (op,) + tuple(id(a) if isinstance(a, Expr) else (a.cons_key() if hasattr(a, "cons_key") else a) for a in args)
        hit = cls._table.get(key)
        if hit is not None:
            return hit
        self = object.__new__(cls)
        self.op, self.args, self._hash = op, args, hash(key)
        cls._table[key] = self
        return self

    def __hash__(self):
        return self._hash

    def __eq__(self, other):
        return self is other

    # ---- operators (ministark: impl Add/Sub/Mul/Div/Neg for Expr, Pow) ------------------------
    @staticmethod
    def _lift(v):
        return v if isinstance(v, Expr) else Constant(v)

    def __add__(self, o): return Expr("add", self, Expr._lift(o))
    def __radd__(self, o): return Expr("add", Expr._lift(o), self)
    def __sub__(self, o): return Expr("sub", self, Expr._lift(o))
    def __rsub__(self, o): return Expr("sub", Expr._lift(o), self)
    def __mul__(self, o): return Expr("mul", self, Expr._lift(o))
    def __rmul__(self, o): return Expr("mul", Expr._lift(o), self)
    def __truediv__(self, o): return Expr("div", self, Expr._lift(o))
    def __rtruediv__(self, o): return Expr("div", Expr._lift(o), self)
    def __neg__(self): return Expr("neg", self)

    def pow(self, e: int):
        if e < 0:
            return Constant(1) / self.pow(-e)
        return Expr("pow", self, int(e))

    def __pow__(self, e): return self.pow(e)

    # AlgebraicItem::Trace helpers used all over the AIR files: `.curr()`, `.next()`, `.offset(k)`
    # are defined on the column enums there; here Trace(col, off) is the primitive.

    def __repr__(self):
        return f"Expr({self.op}, {', '.join(map(repr, self.args))})" if self.op in ("const", "trace", "x", "challenge", "hint") else f"<{self.op}>"


X = Expr("x")


def Constant(v: int) -> Expr:
    return Expr("const", int(v) % P)


def Trace(col: int, offset: int = 0) -> Expr:
    """AlgebraicItem::Trace(column, row_offset) — value of column `col` at row + offset (wrapping)."""
    return Expr("trace", int(col), int(offset))


def Challenge(i: int) -> Expr:
    return Expr("challenge", int(i))


def Hint(i: int) -> Expr:
    return Expr("hint", int(i))


def Periodic(coeffs, interval_size: int) -> Expr:
    """ministark PeriodicColumn::new(&COEFFS, INTERVAL_SIZE) (layouts/src/recursive/air.rs:38-50):
    evaluates to  sum_k coeffs[k] * (X^(n / interval_size))^k."""
    return Expr("periodic", tuple(int(c) % P for c in coeffs), int(interval_size))


def composition_constraint(constraints, coeff_index: int = 0) -> Expr:
    """AirConfig::composition_constraint (layouts/src/recursive/air.rs:1184-1200):
    sum_i constraint_i * CompositionCoeff(0)^i."""
    alpha = Expr("composition_coeff", int(coeff_index))
    total = None
    for i, c in enumerate(constraints):
        term = c * alpha.pow(i)
        total = term if total is None else total + term
    return total
