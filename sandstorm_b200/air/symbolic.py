"""Tracked field elements for the two-phase constraint compiler.

The STRUCTURE of a compiled constraint program (code words, taps, table shapes) does not depend on the
verifier's challenges, the hints or the composition coefficient — only the VALUES of its constants do
(program.py: `structure_hash` is the same for every draw).  The values, however, only exist after the
extension-trace commitment, in the middle of a prove.  So the compiler runs once, ahead of time, on
`Sym` numbers: a `Sym` behaves like a canonical integer mod P (it carries the value of a generic
reference draw, which is what every structural decision looks at) and records the arithmetic applied
to it on a tape.  Re-playing the tape with the real challenges (`Tape.replay`) yields every constant of
the program in a few milliseconds — the per-proof "value patch" that sits on the critical path
between the extension-trace commitment and constraint evaluation (ministark does the equivalent
substitution when it evaluates `Expr` leaves `Challenge(i)` / `Hint(i)` / `CompositionCoeff(i)`)."""
from __future__ import annotations

from .expr import P


class Tape:
    """Straight-line arithmetic over the inputs.  ops[k] = (op, a, b): operands are Sym indices (>= 0, tagged as
    ('s', idx)) or plain ints."""

    _uid = 0

    def __init__(self):
        self.ops: list[tuple] = []
        self.n_inputs = 0
        Tape._uid += 1
        self.uid = Tape._uid

    def input(self, value: int) -> "Sym":
        assert len(self.ops) == self.n_inputs, "inputs first"
        self.ops.append(("in", self.n_inputs, None))
        self.n_inputs += 1
        return Sym(value % P, len(self.ops) - 1, self)

    def push(self, op, a, b, value) -> "Sym":
        self.ops.append((op, a, b))
        return Sym(value, len(self.ops) - 1, self)

    def compact(self, roots: list) -> tuple[list, list]:
        """Keeps only the operations the `roots` (Sym or int) depend on.  Returns (ops, root references) where a
        reference is an int constant or ('s', new index)."""
        need = set()
        stack = [r.i for r in roots if isinstance(r, Sym)]
        while stack:
            k = stack.pop()
            if k in need:
                continue
            need.add(k)
            _, a, b = self.ops[k]
            for x in (a, b):
                if isinstance(x, tuple):
                    stack.append(x[1])
        remap, ops = {}, []
        for k in sorted(need):
            op, a, b = self.ops[k]
            fix = lambda x: ("s", remap[x[1]]) if isinstance(x, tuple) else x
            remap[k] = len(ops)
            ops.append((op, fix(a), fix(b)))
        return ops, [("s", remap[r.i]) if isinstance(r, Sym) else int(r) % P for r in roots]


def replay(ops: list, inputs: list) -> list:
    """values of every tape slot for the given inputs (canonical ints)."""
    vals = [0] * len(ops)
    for k, (op, a, b) in enumerate(ops):
        if op == "in":
            vals[k] = inputs[a] % P
            continue
        x = vals[a[1]] if isinstance(a, tuple) else a
        y = vals[b[1]] if isinstance(b, tuple) else b
        if op == "add": v = x + y
        elif op == "sub": v = x - y
        elif op == "mul": v = x * y
        elif op == "neg": v = -x
        elif op == "pow": v = pow(x, y, P)            # y: plain int exponent (may be negative: inverse power)
        else: raise ValueError(op)
        vals[k] = v % P
    return vals


class Sym:
    """A field element mod P whose value (under the reference draw) drives the compiler's decisions and whose
    derivation is recorded on the tape.  Hashes and compares by value, so it can key the compiler's tables."""
    __slots__ = ("v", "i", "t")

    def __init__(self, v: int, i: int, t: Tape):
        self.v, self.i, self.t = v, i, t

    # -- value semantics ------------------------------------------------------------------------------------
    def __hash__(self): return hash(self.v)
    def __eq__(self, o): return self.v == (o.v if isinstance(o, Sym) else o)
    def __ne__(self, o): return not self.__eq__(o)
    def __lt__(self, o): return self.v < (o.v if isinstance(o, Sym) else o)
    def __le__(self, o): return self.v <= (o.v if isinstance(o, Sym) else o)
    def __gt__(self, o): return self.v > (o.v if isinstance(o, Sym) else o)
    def __ge__(self, o): return self.v >= (o.v if isinstance(o, Sym) else o)
    def __bool__(self): return self.v != 0
    def __repr__(self): return f"Sym({self.v:#x}@{self.i})"
    def cons_key(self): return ("sym", self.t.uid, self.v)

    def __mod__(self, m):
        assert m == P
        return self

    # -- arithmetic ----------------------------------------------------------------------------------------------
    @staticmethod
    def _ref(x):
        return ("s", x.i) if isinstance(x, Sym) else int(x) % P

    @staticmethod
    def _val(x):
        return x.v if isinstance(x, Sym) else int(x) % P

    def _bin(self, op, a, b):
        va, vb = Sym._val(a), Sym._val(b)
        v = {"add": va + vb, "sub": va - vb, "mul": va * vb}[op] % P
        return self.t.push(op, Sym._ref(a), Sym._ref(b), v)

    def __add__(self, o):
        if not isinstance(o, Sym) and o % P == 0:
            return self
        return self._bin("add", self, o)

    __radd__ = __add__

    def __sub__(self, o):
        if isinstance(o, Sym):
            if o.i == self.i:
                return 0
        elif o % P == 0:
            return self
        return self._bin("sub", self, o)

    def __rsub__(self, o):
        return self._bin("sub", o, self)

    def __mul__(self, o):
        if not isinstance(o, Sym):
            o %= P
            if o == 0:
                return 0
            if o == 1:
                return self
        return self._bin("mul", self, o)

    __rmul__ = __mul__

    def __neg__(self):
        return self.t.push("neg", ("s", self.i), None, -self.v % P)

    def __pow__(self, e, mod=None):
        assert mod in (None, P) and isinstance(e, int)
        if e == 1:
            return self
        if e == 0:
            return 1
        return self.t.push("pow", ("s", self.i), e, pow(self.v, e, P))


def value_of(x) -> int:
    """the reference value of a Sym, or the int itself."""
    return x.v if isinstance(x, Sym) else x
