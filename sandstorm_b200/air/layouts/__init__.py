"""Symbolic AIRs of the reference's layouts (plain / recursive / starknet), stored as data.

`<layout>.json` is generated from `layouts/src/<layout>/air.rs` by tools/air_transpile.py (the
reference builds the same DAG at run time in `AirConfig::constraints(trace_len)`; e.g.
layouts/src/recursive/air.rs:61-1182).  `Layout.constraints(n)` instantiates it for a trace length
and returns `sandstorm_b200.air.Expr` trees; `Layout.composition(n)` is
`AirConfig::composition_constraint` (air.rs:1184-1200): sum_i constraint_i * alpha^i."""
from __future__ import annotations

import json
import os
from functools import lru_cache

from ..expr import Challenge, Constant, Expr, Hint, Periodic, Trace, X, composition_constraint, P

_HERE = os.path.dirname(os.path.abspath(__file__))


@lru_cache(maxsize=None)
def _periodic_coeffs() -> dict:
    with open(os.path.join(_HERE, "periodic_coeffs.json")) as f:
        return {k: [int(v, 16) for v in vals] for k, vals in json.load(f).items()}


class Layout:
    def __init__(self, name: str):
        with open(os.path.join(_HERE, f"{name}.json")) as f:
            d = json.load(f)
        self.name = name
        self.num_base_columns = d["num_base_columns"]
        self.num_extension_columns = d["num_extension_columns"]
        self.cycle_height = d["cycle_height"]
        self.n_constraints = d["n_constraints"]
        self.max_offset = d["max_offset"]
        self._nodes, self._constraints, self._periodic = d["nodes"], d["constraints"], d["periodic"]

    @property
    def num_columns(self) -> int:
        return self.num_base_columns + self.num_extension_columns

    def taps(self) -> list[tuple[int, int]]:
        """air.trace_arguments(): the distinct (column, row offset) cells the constraints read."""
        return sorted({(k[1], k[2]) for k in self._nodes if k[0] == "trace"})

    def min_trace_len(self) -> int:
        m = self.cycle_height
        for spec in self._periodic.values():
            m = max(m, spec["interval"])
        return m

    def constraints(self, n: int, inv_x_minus_one_col: int | None = None) -> list[Expr]:
        """inv_x_minus_one_col: index of an auxiliary matrix column holding w[i] = 1 / (x_i - 1)
        (ss_inv_x_minus_c with c = 1).  Divisions by the boundary terms X - g^e are then rewritten as
        g^-e * w[i - blowup * e] — a shifted read instead of a per-row inversion
        (x_i - g^e = g^e (x_{i - b e} - 1))."""
        if n & (n - 1) or n < self.min_trace_len():
            raise ValueError(f"trace length {n} is not a power of two >= {self.min_trace_len()}")
        g = pow(3, (P - 1) // n, P)
        coeffs = _periodic_coeffs()
        built: list[Expr] = []
        nodes = self._nodes

        def boundary_exponent(idx):
            """e if node idx is X - g^e, else None."""
            k = nodes[idx]
            if k[0] != "sub" or nodes[k[1]][0] != "x":
                return None
            c = nodes[k[2]]
            if c[0] == "gpow":
                return (c[1] * n // c[2] + c[3]) % n
            if c[0] == "const" and int(c[1], 16) == 1:
                return 0
            return None

        for k in nodes:
            op = k[0]
            if op == "div" and inv_x_minus_one_col is not None:
                e_b = boundary_exponent(k[2])
                if e_b is not None:
                    built.append(built[k[1]] * Constant(pow(g, -e_b, P)) * Trace(inv_x_minus_one_col, -e_b))
                    continue
            if op == "x": e = X
            elif op == "const": e = Constant(int(k[1], 16))
            elif op == "gpow": e = Constant(pow(g, (k[1] * n // k[2] + k[3]) % n, P))
            elif op == "xpow": e = X.pow(k[1] * n // k[2] + k[3])
            elif op == "trace": e = Trace(k[1], k[2])
            elif op == "challenge": e = Challenge(k[1])
            elif op == "hint": e = Hint(k[1])
            elif op == "periodic": e = Periodic(coeffs[k[1]], k[2])
            elif op == "neg": e = -built[k[1]]
            elif op == "pow": e = built[k[1]].pow(k[2])
            elif op == "add": e = built[k[1]] + built[k[2]]
            elif op == "sub": e = built[k[1]] - built[k[2]]
            elif op == "mul": e = built[k[1]] * built[k[2]]
            elif op == "div": e = built[k[1]] / built[k[2]]
            else: raise ValueError(op)
            built.append(e)
        return [built[i] for i in self._constraints]

    def composition(self, n: int, inv_x_minus_one_col: int | None = None) -> Expr:
        return composition_constraint(self.constraints(n, inv_x_minus_one_col))

    def n_challenges(self) -> int:
        return 1 + max((k[1] for k in self._nodes if k[0] == "challenge"), default=-1)

    def n_hints(self) -> int:
        return 1 + max((k[1] for k in self._nodes if k[0] == "hint"), default=-1)


@lru_cache(maxsize=None)
def load_layout(name: str) -> Layout:
    return Layout(name)
