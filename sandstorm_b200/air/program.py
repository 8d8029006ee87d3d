"""Compiler: composition-constraint Expr DAG -> straight-line program for `ss_constraint_eval`.

What the reference does at run time with an expression interpreter over whole vectors (ministark's
evaluator) is split here between host and device:

  host   : substitute challenges / hints / composition coefficient, fold constants, hash-cons (CSE),
           classify every node by its PERIOD in the LDE row index i (x_i = 3 * w_N^i):
             period 1         -> constant
             period T <= 2^17 -> lookup table of T entries (zerofiers X^(n/k) - c and their inverses,
                                 periodic columns P(X^(n/interval)), products of those)
             full             -> depends on trace cells or on X itself -> device instruction
           then rewrite the full-period part as LINEAR COMBINATIONS over non-linear atoms:
             * constant factors are folded into coefficients at compile time;
             * the root sum  sum_i alpha^i * num_i * D_i  is regrouped by shared factors D (zerofier
               tables, boundary-denominator taps):  sum_D D * (sum_{i in D} alpha^i * num_i);
             * a linear combination with general coefficients becomes one DOT instruction: the products
               are accumulated unreduced (512 bits) and Montgomery-reduced once;
             * every value carries an exact upper bound, so additions are raw, subtractions add a
               pre-computed multiple of p, and a reduction (RED) is emitted only where a bound would
               overflow 2^256 — the device never compares magnitudes.
  device : one thread per LDE row runs the instruction list over a small slot file (liveness-allocated);
           constants, table entries and trace taps are fetched directly as instruction operands.

Program blob, version 2 (little-endian u32 words unless noted):
  [0] magic 'SSCP'  [1] version  [2] n_words (16-byte code words)  [3] n_consts  [4] n_tables  [5] n_slots
  [6] log_n (trace) [7] log_blowup  [8] n_taps  [9..15] reserved
  then n_tables x (log_period | (scale + 1) << 8, offset in elements), padded to an even count
       scale = index of a constant the stored table has to be multiplied by before use (ss_constraint_eval does it on
       the device when it uploads the blob): a table that is a challenge-dependent coefficient times a
       challenge-independent periodic function keeps its precomputed values, and only the coefficient is patched
  then n_taps x (column, row offset mod N in LDE rows), padded to an even count
  then n_words x 4 words of code     then consts (32 B each)   then table data (32 B each)
Code word:  (op | dst << 8 | n << 16, A, B, 0);  a DOT is followed by ceil(n / 2) words (A0, B0, A1, B1).
Operand word:  kind << 29 | payload   kind 0 slot | 1 const index | 2 tap index | 3 table index | 4 x_i.
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

from .expr import Expr, P
from .symbolic import Sym, Tape, replay, value_of

R = 2**256
MAGIC = 0x50435353          # 'SSCP'
VERSION = 3
MAX_TABLE_LOG = 17
GENERATOR = 3

OP_NOP, OP_MOV, OP_ADD, OP_SUBK, OP_RED, OP_MUL, OP_DOT, OP_INV, OP_OUT = range(9)
K_SLOT, K_CONST, K_TAP, K_TABLE, K_X = range(5)
FULL = 0    # period marker for row-dependent nodes


def _mont_limbs(v: int) -> list[int]:
    m = v % P * R % P
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def _ntt(a: list, root: int) -> list:
    """plain radix-2 transform: out[j] = sum_m a[m] * root^(j m), len(a) a power of two, root of that order."""
    n = len(a)
    if n == 1:
        return a
    bits = n.bit_length() - 1
    a = [a[int(f"{i:0{bits}b}"[::-1], 2)] for i in range(n)]
    size = 2
    while size <= n:
        wn, half = pow(root, n // size, P), size // 2
        tw, t = [1] * half, 1
        for j in range(half):
            tw[j] = t
            t = t * wn % P
        for start in range(0, n, size):
            for j in range(half):
                u, v = a[start + j], a[start + j + half] * tw[j] % P
                a[start + j], a[start + j + half] = (u + v) % P, (u - v) % P
        size *= 2
    return a


@dataclass
class CompiledProgram:
    blob: bytes
    n_instr: int
    n_consts: int
    n_tables: int
    n_slots: int
    n_mul: int
    n_addsub: int
    n_trace_taps: int
    n_batch_inv: int
    table_sizes: list = field(default_factory=list)
    n_red: int = 0
    n_dot: int = 0
    n_inv: int = 0


@dataclass
class ProgramTemplate:
    """A constraint program compiled ahead of the proof on tracked challenges (air/symbolic.py): everything but the
    values of the challenge-dependent constants.  `patch` substitutes the real challenges / hints / composition
    coefficients — a tape replay over a few thousand field operations plus one copy of the blob — and is the only
    part of the compilation on the critical path of a prove (between the extension-trace commitment and constraint
    evaluation).  The result is bit-identical to compiling with the values directly: same code, same constants."""
    head: bytes                      # header, table descriptors, taps, code (padded to 32 bytes)
    const_refs: list                 # per constant: canonical int or ('s', tape slot)
    tape_ops: list
    n_inputs: tuple                  # (n_challenges, n_hints, n_coeffs)
    table_bytes: bytes
    stats: "CompiledProgram"

    def patch(self, challenges=(), hints=(), composition_coeffs=(0,)) -> CompiledProgram:
        nc, nh, nk = self.n_inputs
        inputs = list(challenges) + list(hints) + list(composition_coeffs)
        if (len(challenges), len(hints), len(composition_coeffs)) != (nc, nh, nk):
            raise ValueError(f"template expects {nc} challenges, {nh} hints, {nk} composition coefficients")
        vals = replay(self.tape_ops, inputs)
        consts = [vals[r[1]] if isinstance(r, tuple) else r for r in self.const_refs]
        body = np.array([l for v in consts for l in _mont_limbs(v)], dtype=np.uint64).tobytes() if consts else b""
        st = self.stats
        return CompiledProgram(blob=b"".join((self.head, body, self.table_bytes)), n_instr=st.n_instr, n_consts=st.n_consts,
                               n_tables=st.n_tables, n_slots=st.n_slots, n_mul=st.n_mul, n_addsub=st.n_addsub,
                               n_trace_taps=st.n_trace_taps, n_batch_inv=st.n_batch_inv, table_sizes=st.table_sizes, n_red=st.n_red,
                               n_dot=st.n_dot, n_inv=st.n_inv)


class _Lower:
    """Lowered IR node: ('const', v) | ('table', values) | ('x',) | ('trace', col, off) |
    ('add'|'sub'|'mul', a, b) | ('neg', a) | ('inv', a)"""

    def __init__(self, log_n, log_blowup, challenges, hints, coeffs):
        self.log_n, self.log_b = log_n, log_blowup
        self.n, self.N = 1 << log_n, 1 << (log_n + log_blowup)
        self.challenges, self.hints, self.coeffs = challenges, hints, coeffs
        self.w = pow(GENERATOR, (P - 1) // self.N, P)
        self.memo_period: dict = {}
        self.memo_norm: dict = {}
        self.memo_eval: dict = {}

    # -- 1. normalisation: substitute symbols, expand pow of non-X, fold constants -------------------
    def norm(self, e: Expr) -> Expr:
        hit = self.memo_norm.get(e)
        if hit is not None:
            return hit
        op = e.op
        if op in ("x", "const", "trace", "periodic"):
            r = e
        elif op == "challenge":
            r = Expr("const", self.challenges[e.args[0]] % P)
        elif op == "hint":
            r = Expr("const", self.hints[e.args[0]] % P)
        elif op == "composition_coeff":
            r = Expr("const", self.coeffs[e.args[0]] % P)
        elif op == "pow":
            base, k = self.norm(e.args[0]), e.args[1]
            if base.op == "const":
                r = Expr("const", pow(base.args[0], k, P))
            elif base.op == "x":
                r = Expr("const", 1) if k == 0 else (base if k == 1 else Expr("pow", base, k))
            else:
                acc, sq, kk = None, base, k
                while kk:
                    if kk & 1:
                        acc = sq if acc is None else self.norm(Expr("mul", acc, sq))
                    kk >>= 1
                    if kk:
                        sq = self.norm(Expr("mul", sq, sq))
                r = acc if acc is not None else Expr("const", 1)
        else:
            args = tuple(self.norm(a) for a in e.args)
            if all(a.op == "const" for a in args):
                vals = [a.args[0] for a in args]
                if op == "add": v = vals[0] + vals[1]
                elif op == "sub": v = vals[0] - vals[1]
                elif op == "mul": v = vals[0] * vals[1]
                elif op == "neg": v = -vals[0]
                elif op == "div": v = vals[0] * pow(vals[1], -1, P)
                else: raise ValueError(op)
                r = Expr("const", v % P)
            else:
                r = Expr(op, *args)
        self.memo_norm[e] = r
        return r

    # -- 2. period of a normalised node ------------------------------------------------------------
    def period(self, e: Expr) -> int:
        hit = self.memo_period.get(e)
        if hit is not None:
            return hit
        op = e.op
        if op == "const":
            p = 1
        elif op in ("x", "trace"):
            p = FULL
        elif op == "pow":                                       # X^k
            p = self.N // math.gcd(self.N, e.args[1])
        elif op == "periodic":
            interval = e.args[1]
            if interval > self.n or self.n % interval:
                raise ValueError("periodic column interval must divide the trace length")
            p = self.N // math.gcd(self.N, self.n // interval)
        else:
            ps = [self.period(a) for a in e.args]
            p = FULL if any(q == FULL for q in ps) else max(ps)
        if p != FULL and p > (1 << MAX_TABLE_LOG):
            p = FULL
        self.memo_period[e] = p
        return p

    # -- 3. big-int evaluation of a periodic node at row j ----------------------------------------------
    def eval_at(self, e: Expr, j: int) -> int:
        key = (e, j)
        hit = self.memo_eval.get(key)
        if hit is not None:
            return hit
        op = e.op
        if op == "const":
            v = e.args[0]
        elif op == "pow":
            k = e.args[1]
            v = pow(GENERATOR, k, P) * pow(self.w, (k * j) % self.N, P) % P
        elif op == "x":
            v = GENERATOR * pow(self.w, j % self.N, P) % P
        elif op == "periodic":
            coeffs, interval = e.args
            k = self.n // interval
            y = pow(GENERATOR, k, P) * pow(self.w, (k * j) % self.N, P) % P
            v = 0
            for c in reversed(coeffs):
                v = (v * y + c) % P
        else:
            a = [self.eval_at(x, j) for x in e.args]
            if op == "add": v = (a[0] + a[1]) % P
            elif op == "sub": v = (a[0] - a[1]) % P
            elif op == "mul": v = a[0] * a[1] % P
            elif op == "neg": v = -a[0] % P
            elif op == "div":
                if a[1] == 0:
                    raise ZeroDivisionError("periodic denominator vanishes on the LDE coset")
                v = a[0] * pow(a[1], -1, P) % P
            else:
                raise ValueError(op)
        self.memo_eval[key] = v
        return v

    # -- 4. all values of a periodic node over one period (rows 0 .. period-1), vectorised over the rows ----------
    def table_values(self, e: Expr, memo: dict | None = None) -> list:
        memo = {} if memo is None else memo
        hit = memo.get(e)
        if hit is not None:
            return hit
        op = e.op
        T = self.period(e)
        assert T != FULL

        def powers(k: int, count: int) -> list:
            """(3 * w^j)^k for j < count."""
            step, v, out = pow(self.w, k % self.N, P), pow(GENERATOR, k, P), []
            for _ in range(count):
                out.append(v)
                v = v * step % P
            return out

        if op == "const":
            vals = [e.args[0] % P]
        elif op == "pow":
            vals = powers(e.args[1], T)
        elif op == "x":
            vals = powers(1, T)
        elif op == "periodic":
            coeffs, interval = e.args
            k = self.n // interval
            if len(coeffs) <= 16 or len(coeffs) > T:
                ys = powers(k, T)
                vals = [0] * T
                for c in reversed(coeffs):
                    vals = [(v * y + c) % P for v, y in zip(vals, ys)]
            else:
                # the T points are y_j = 3^k * W^j with W = w^k of order T: one size-T transform of the coefficients
                # scaled by (3^k)^m instead of len(coeffs) * T Horner steps
                shift, acc, a = pow(GENERATOR, k, P), 1, []
                for c in coeffs:
                    a.append(c * acc % P)
                    acc = acc * shift % P
                vals = _ntt(a + [0] * (T - len(a)), pow(self.w, k % self.N, P))
        else:
            args = [self.table_values(a, memo) for a in e.args]
            a = args[0]
            if len(a) != T:
                a = a * (T // len(a))
            if len(args) > 1:
                b = args[1]
                if len(b) != T:
                    b = b * (T // len(b))
            if op == "add": vals = [(x + y) % P for x, y in zip(a, b)]
            elif op == "sub": vals = [(x - y) % P for x, y in zip(a, b)]
            elif op == "mul": vals = [x * y % P for x, y in zip(a, b)]
            elif op == "neg": vals = [-x % P for x in a]
            elif op == "div":
                # Montgomery's trick: one modular inversion per table
                pre, acc = [], 1
                for y in b:
                    if y == 0:
                        raise ZeroDivisionError("periodic denominator vanishes on the LDE coset")
                    pre.append(acc)
                    acc = acc * y % P
                inv = pow(acc, -1, P)
                invs = [0] * T
                for j in range(T - 1, -1, -1):
                    invs[j] = inv * pre[j] % P
                    inv = inv * b[j] % P
                vals = [x * y % P for x, y in zip(a, invs)]
            else:
                raise ValueError(op)
        memo[e] = vals
        return vals


LIM = 1 << 256                       # every stored value must stay below this
RED_BOUND = 1 << 252                 # bound after RED
OUT_CAP = 8 * P                      # products are kept below this so that sums of a few of them still fit
ACC_LIM = (1 << 512) - P * R         # an unreduced accumulator (plus the p * 2^256 bias) must fit 512 bits
ACC_SOFT = 7 * P * R                 # ... and its reduction should come out below OUT_CAP
DOT_SLOT_CHUNK = 8                   # slot operands alive at once inside one DOT
DOT_MAX_TERMS = 4096


SMALL_COEFF = 16                     # |c| <= this: c * A is a short chain of additions, not a multiplication


def _small(c: int) -> int:
    """signed value of a coefficient when it is a small integer (2 <= |s| <= SMALL_COEFF), else 0."""
    c = value_of(c)
    if 2 <= c <= SMALL_COEFF:
        return c
    if 2 <= P - c <= SMALL_COEFF:
        return c - P
    return 0


def _mul_bound(ba: int, bb: int) -> int:
    """exclusive upper bound of fp::mul / acc_reduce for operands below ba, bb (fp252.cuh: result <= a*b/R + p)."""
    return ba * bb // R + P + 1


class _Graph:
    """Hash-consed device DAG.  kinds: const(v) table(tid) x trace(col, off_lde) mul(a, b) inv(a) lc(((node, coeff), ...), c0)"""

    def __init__(self):
        self.kind: list[str] = []
        self.args: list[tuple] = []
        self.index: dict = {}

    def mk(self, kind: str, *args) -> int:
        key = (kind,) + args
        hit = self.index.get(key)
        if hit is None:
            hit = len(self.kind)
            self.kind.append(kind)
            self.args.append(args)
            self.index[key] = hit
        return hit

    def children(self, nid: int):
        k, a = self.kind[nid], self.args[nid]
        if k == "mul":
            return list(a)
        if k == "inv":
            return [a[0]]
        if k == "lc":
            return [t for t, _ in a[0]]
        return []


LEAF_KINDS = ("const", "table", "trace")


def compile_template(expr: Expr, log_n: int, log_blowup: int, n_challenges: int = 0, n_hints: int = 0, n_coeffs: int = 1,
                     max_slots: int = 256, regroup: bool = True, with_tables: bool = True, seed: int = 0x5EED) -> ProgramTemplate:
    """Compiles `expr` with the challenges / hints / composition coefficients left open (see ProgramTemplate).  The
    reference draw that steers the structural decisions is generic (seeded random field elements), so the structure is
    the one every real draw has, including draws whose hints happen to be 0 or 1."""
    import random

    rnd, tape = random.Random(seed), Tape()
    syms = [tape.input(rnd.randrange(2, P)) for _ in range(n_challenges + n_hints + n_coeffs)]
    return compile_program(expr, log_n, log_blowup, syms[:n_challenges], syms[n_challenges:n_challenges + n_hints],
                           syms[n_challenges + n_hints:], max_slots, regroup, with_tables, _tape=tape)


def compile_program(expr: Expr, log_n: int, log_blowup: int, challenges=(), hints=(), composition_coeffs=(0,),
                    max_slots: int = 256, regroup: bool = True, with_tables: bool = True, _tape: Tape | None = None):
    """expr: the composition constraint (or any Expr).  challenges / hints / composition_coeffs: canonical ints.
    (With _tape — compile_template — they are tracked symbols and the result is a ProgramTemplate.)"""
    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 200000))
    lw = _Lower(log_n, log_blowup, list(challenges), list(hints), list(composition_coeffs))
    root = lw.norm(expr)
    g = _Graph()

    consts: list[int] = []
    const_ix: dict[int, int] = {}
    tap_list: list[tuple] = []           # distinct (column, row offset mod N)
    tap_ix: dict[tuple, int] = {}
    tables: list = []                    # Expr per table (values are produced at serialisation time)
    table_ix: dict[Expr, int] = {}

    def const_id(v: int) -> int:
        v %= P
        if v not in const_ix:
            const_ix[v] = len(consts)
            consts.append(v)
        return const_ix[v]

    def table_node(e: Expr) -> int:
        if e not in table_ix:
            table_ix[e] = len(tables)
            tables.append(e)
        return g.mk("table", table_ix[e])

    # ---- 1. linear combinations over atoms ---------------------------------------------------------------
    def node_of(lc) -> int:
        terms, c0 = lc
        if not terms:
            return g.mk("const", c0 % P)
        if len(terms) == 1 and c0 == 0:
            (nid, c), = terms.items()
            if c == 1:
                return nid
        return g.mk("lc", tuple(sorted(terms.items())), c0 % P)

    def lc_add(a, b, sign=1):
        terms = dict(a[0])
        for k, v in b[0].items():
            nv = (terms.get(k, 0) + sign * v) % P
            if nv:
                terms[k] = nv
            else:
                terms.pop(k, None)
        return terms, (a[1] + sign * b[1]) % P

    def lc_scale(a, c):
        c %= P
        if c == 0:
            return {}, 0
        if c == 1:
            return a
        terms, c0 = a
        if c == P - 1 or len(terms) <= 1 or all(v not in (1, P - 1) and not _small(v) for v in terms.values()):
            return {k: v * c % P for k, v in terms.items()}, c0 * c % P
        return {node_of(a): c}, 0

    def lc_split(a):
        """(coefficient, node) with a == coefficient * node."""
        terms, c0 = a
        if len(terms) == 1 and c0 == 0:
            (nid, c), = terms.items()
            return c, nid
        return 1, node_of(a)

    def lc_mul(a, b):
        if not a[0]:
            return lc_scale(b, a[1])
        if not b[0]:
            return lc_scale(a, b[1])
        ca, na = lc_split(a)
        cb, nb = lc_split(b)
        return {g.mk("mul", min(na, nb), max(na, nb)): ca * cb % P}, 0

    memo_lc: dict[Expr, tuple] = {}
    ONE = Expr("const", 1)

    def lower(e: Expr):
        hit = memo_lc.get(e)
        if hit is not None:
            return hit
        p = lw.period(e)
        if p == 1:
            r = ({}, lw.eval_at(e, 0) % P)
        elif p != FULL:
            r = ({table_node(e): 1}, 0)
        elif e.op == "x":
            r = ({g.mk("x"): 1}, 0)
        elif e.op == "trace":
            r = ({g.mk("trace", e.args[0], e.args[1] * (1 << log_blowup)): 1}, 0)
        elif e.op == "pow":                                      # X^k with a long period: square-and-multiply on X
            k, acc, sq = e.args[1], None, ({g.mk("x"): 1}, 0)
            while k:
                if k & 1:
                    acc = sq if acc is None else lc_mul(acc, sq)
                k >>= 1
                if k:
                    sq = lc_mul(sq, sq)
            r = acc
        elif e.op == "periodic":
            raise ValueError("periodic column with a period above the table limit")
        elif e.op == "div":
            num, den = e.args
            if lw.period(den) != FULL:
                r = lc_mul(lower(num), lower(Expr("div", ONE, den)))
            else:
                r = lc_mul(lower(num), ({g.mk("inv", node_of(lower(den))): 1}, 0))
        elif e.op == "neg":
            r = lc_scale(lower(e.args[0]), P - 1)
        elif e.op == "add":
            r = lc_add(lower(e.args[0]), lower(e.args[1]))
        elif e.op == "sub":
            r = lc_add(lower(e.args[0]), lower(e.args[1]), -1)
        elif e.op == "mul":
            r = lc_mul(lower(e.args[0]), lower(e.args[1]))
        else:
            raise ValueError(e.op)
        memo_lc[e] = r
        return r

    root_lc = lower(root)

    # ---- 2. regroup the root sum by shared factors -----------------------------------------------------------
    def flatten(nid: int) -> list[int]:
        if g.kind[nid] == "mul":
            return flatten(g.args[nid][0]) + flatten(g.args[nid][1])
        return [nid]

    def product(ids: list[int]) -> int:
        """one node for the product of `ids`: periodic factors are merged into a single table first."""
        tabs = [i for i in ids if g.kind[i] == "table"]
        rest = sorted(i for i in ids if g.kind[i] != "table")
        if len(tabs) > 1:
            e = tables[g.args[tabs[0]][0]]
            for t in tabs[1:]:
                e = Expr("mul", e, tables[g.args[t][0]])
            tabs = [table_node(e)]
        acc = None
        for i in tabs + rest:
            acc = i if acc is None else g.mk("mul", min(acc, i), max(acc, i))
        return acc

    if regroup and len(root_lc[0]) > 1:
        facs = {nid: flatten(nid) for nid in root_lc[0]}
        usage: dict[int, int] = {}
        for fl in facs.values():
            for f in set(fl):
                usage[f] = usage.get(f, 0) + 1
        groups: dict[tuple, list] = {}
        new_terms: dict[int, int] = {}

        def add_term(nid, c):
            nv = (new_terms.get(nid, 0) + c) % P
            if nv:
                new_terms[nid] = nv
            else:
                new_terms.pop(nid, None)

        for nid, c in root_lc[0].items():
            fl = facs[nid]
            shared = [f for f in fl if g.kind[f] in ("table", "inv") or usage[f] > 1]
            own = [f for f in fl if not (g.kind[f] in ("table", "inv") or usage[f] > 1)]
            if not own:                                       # everything is shared: keep the least shared factor as the numerator
                k = min(range(len(shared)), key=lambda j: (usage[shared[j]], g.kind[shared[j]] in LEAF_KINDS))
                own, shared = [shared[k]], shared[:k] + shared[k + 1:]
            if not shared:
                add_term(nid, c)
                continue
            groups.setdefault(tuple(sorted(shared)), []).append((product(own), c))
        for key, members in groups.items():
            key = list(key)
            inner: dict[int, int] = {}
            for nid, c in members:
                nv = (inner.get(nid, 0) + c) % P
                if nv:
                    inner[nid] = nv
                else:
                    inner.pop(nid, None)
            if not inner:
                continue
            if len(inner) == 1:
                (nid, c), = inner.items()
                tabs = [f for f in key if g.kind[f] == "table"]
                if c != 1 and tabs:                            # fold the coefficient into the table
                    e = Expr("mul", Expr("const", c), tables[g.args[tabs[0]][0]])
                    key[key.index(tabs[0])] = table_node(e)
                    add_term(product([nid] + key), 1)
                elif c != 1:
                    add_term(product([nid] + key), c)
                else:
                    add_term(product([nid] + key), 1)
            else:
                add_term(product([node_of((inner, 0))] + key), 1)
        root_lc = (new_terms, root_lc[1])
    root_node = node_of(root_lc)

    # ---- 3. batched inversion of the independent full-period denominators (Montgomery's trick as DAG nodes) ---
    reach: set[int] = set()
    order: list[int] = []

    def visit(nid):
        if nid in reach:
            return
        reach.add(nid)
        for ch in g.children(nid):
            visit(ch)
        order.append(nid)

    visit(root_node)
    has_inv: dict[int, bool] = {}
    for nid in order:                                          # children before parents
        has_inv[nid] = g.kind[nid] == "inv" or any(has_inv[ch] for ch in g.children(nid))
    inv_nodes = [nid for nid in order if g.kind[nid] == "inv" and not has_inv[g.args[nid][0]]]
    subst: dict[int, int] = {}
    if len(inv_nodes) > 1:
        dens = [g.args[nid][0] for nid in inv_nodes]
        pre = [dens[0]]
        for d in dens[1:]:
            pre.append(g.mk("mul", min(pre[-1], d), max(pre[-1], d)))
        run = g.mk("inv", pre[-1])
        for k in range(len(dens) - 1, 0, -1):
            subst[inv_nodes[k]] = g.mk("mul", min(run, pre[k - 1]), max(run, pre[k - 1]))
            run = g.mk("mul", min(run, dens[k]), max(run, dens[k]))
        subst[inv_nodes[0]] = run

    def resolve(nid: int) -> int:
        return subst.get(nid, nid)

    # ---- 4. schedule: liveness-allocated slots, exact bounds, operand-fused three-address code ----------------
    uses: dict[int, int] = {}

    def count(nid):
        nid = resolve(nid)
        uses[nid] = uses.get(nid, 0) + 1
        if uses[nid] == 1:
            for ch in g.children(nid):
                count(ch)

    count(root_node)

    code: list[tuple] = []
    slot_of: dict[int, int] = {}
    bound: dict[int, int] = {}
    free: list[int] = []
    n_slots = 0
    stats = {"mul": 0, "addsub": 0, "trace": 0, "red": 0, "dot": 0, "inv": 0}
    X_BOUND = _mul_bound(P, P)

    def alloc() -> int:
        nonlocal n_slots
        if free:
            return free.pop()
        n_slots += 1
        return n_slots - 1

    class Opnd:
        """an instruction operand: a leaf fetched in place, or a slot (owned by a node or temporary)."""
        __slots__ = ("word", "slot", "nid", "keep")

        def __init__(self, word, slot=None, nid=None, keep=False):
            self.word, self.slot, self.nid, self.keep = word, slot, nid, keep      # keep: an alias, released by its owner

        @property
        def b(self) -> int:
            return bound[self.slot] if self.slot is not None else P

    def leaf_opnd(nid: int) -> Opnd:
        k, a = g.kind[nid], g.args[nid]
        if k == "const":
            return Opnd(K_CONST << 29 | const_id(a[0]))
        if k == "table":
            return Opnd(K_TABLE << 29 | a[0])
        col, off = a
        key = (col, off % lw.N)
        if key not in tap_ix:
            tap_ix[key] = len(tap_list)
            tap_list.append(key)
        stats["trace"] += 1
        return Opnd(K_TAP << 29 | tap_ix[key])

    def operand(nid: int) -> Opnd:
        nid = resolve(nid)
        if g.kind[nid] in LEAF_KINDS:
            return Opnd(leaf_opnd(nid).word, None, nid)
        if nid not in slot_of:
            emit(nid)
        return Opnd(K_SLOT << 29 | slot_of[nid], slot_of[nid], nid)

    def release(op: Opnd):
        if op.keep:
            return
        if op.nid is None:                                      # temporary
            if op.slot is not None:
                free.append(op.slot)
            return
        uses[op.nid] -= 1
        if uses[op.nid] == 0 and op.nid in slot_of:
            free.append(slot_of.pop(op.nid))

    def red(op: Opnd):
        assert op.slot is not None
        code.append((OP_RED | op.slot << 8, op.word, 0, 0))
        bound[op.slot] = RED_BOUND
        stats["red"] += 1

    def reducible(op: Opnd) -> bool:
        return op.slot is not None and bound[op.slot] > RED_BOUND

    def put(opc: int, n: int, a: Opnd, b: Opnd | None, out_bound: int) -> Opnd:
        """emit  tmp = op(a, b); the operands are released first so that the result may reuse one of their slots
        (the device reads both operands completely before it writes the destination)."""
        assert out_bound <= LIM, "bound analysis failed"
        wa, wb = a.word, (b.word if b is not None else 0)
        release(a)
        if b is not None:
            release(b)
        d = alloc()
        assert d < 256 and n < 65536
        code.append((opc | d << 8 | n << 16, wa, wb, 0))
        bound[d] = out_bound
        return Opnd(K_SLOT << 29 | d, d, None)

    def adopt(op: Opnd, nid: int):
        """the temporary `op` becomes the value of node nid (copied first when it belongs to somebody else)."""
        if op.nid is not None or op.slot is None:
            op = put(OP_MOV, 0, op, None, op.b)
        slot_of[nid] = op.slot

    def pick_red(a: Opnd, b: Opnd) -> Opnd:
        if reducible(a) and (a.b >= b.b or not reducible(b)):
            return a
        assert reducible(b), "bound analysis failed"
        return b

    def do_add(a: Opnd, b: Opnd) -> Opnd:
        while a.b + b.b > LIM:
            red(pick_red(a, b))
        stats["addsub"] += 1
        return put(OP_ADD, 0, a, b, a.b + b.b)

    def do_sub(a: Opnd, b: Opnd) -> Opnd:
        while True:
            k = -(-b.b // P)
            if k <= 31 and a.b + k * P <= LIM:
                break
            if reducible(b) and (k > 2 or not reducible(a)):
                red(b)
            else:
                assert reducible(a), "bound analysis failed (sub)"
                red(a)
        stats["addsub"] += 1
        return put(OP_SUBK, k, a, b, a.b + k * P)

    def do_mul(a: Opnd, b: Opnd) -> Opnd:
        while _mul_bound(a.b, b.b) > OUT_CAP and (reducible(a) or reducible(b)):
            red(pick_red(a, b))
        assert a.b * b.b <= ACC_LIM
        stats["mul"] += 1
        return put(OP_MUL, 0, a, b, _mul_bound(a.b, b.b))

    def do_dot(pairs: list) -> Opnd:
        """pairs: [(Opnd, Opnd)]; the caller keeps the sum of bound products within ACC_SOFT."""
        total = sum(a.b * b.b for a, b in pairs)
        assert total <= ACC_LIM and 0 < len(pairs) < 65536
        if len(pairs) == 1:
            return do_mul(pairs[0][0], pairs[0][1])
        words = [w for a, b in pairs for w in (a.word, b.word)]
        if len(pairs) & 1:
            words += [0, 0]
        for a, b in pairs:
            release(a)
            release(b)
        d = alloc()
        assert d < 256
        code.append((OP_DOT | d << 8 | len(pairs) << 16, 0, 0, 0))
        for k in range(0, len(words), 4):
            code.append(tuple(words[k:k + 4]))
        bound[d] = _mul_bound(1, total)
        assert bound[d] <= LIM
        stats["mul"] += len(pairs)
        stats["dot"] += 1
        return Opnd(K_SLOT << 29 | d, d, None)

    def multiple(op: Opnd, m: int) -> Opnd:
        """m * op for a small integer m >= 2 by left-to-right double-and-add on raw additions; releases op."""
        alias = lambda o: Opnd(o.word, o.slot, o.nid, keep=True)
        cur = None
        for bit in bin(m)[3:]:
            src = cur if cur is not None else op
            nxt = do_add(alias(src), alias(src))
            if cur is not None:
                release(cur)
            cur = nxt
            if bit == "1":
                nxt = do_add(alias(cur), alias(op))
                release(cur)
                cur = nxt
            stats["small"] = stats.get("small", 0) + 1
        release(op)
        return cur

    def const_opnd(v: int) -> Opnd:
        return Opnd(K_CONST << 29 | const_id(v))

    def emit_lc(nid: int):
        terms, c0 = g.args[nid]
        small = [(t, _small(c)) for t, c in terms if _small(c)]
        gen = [(t, c) for t, c in terms if c not in (1, P - 1) and not _small(c)]
        plus = [t for t, c in terms if c == 1]
        minus = [t for t, c in terms if c == P - 1]
        gen.sort(key=lambda tc: g.kind[resolve(tc[0])] in LEAF_KINDS)     # slot operands first, leaves fill the chunks up
        acc: Opnd | None = None
        chunk: list = []
        chunk_total = chunk_slots = 0
        for t, c in gen:
            a = operand(t)
            if reducible(a) and a.b > 8 * P:
                red(a)
            w = a.b * P
            if chunk and (chunk_total + w > ACC_SOFT or (a.slot is not None and chunk_slots >= DOT_SLOT_CHUNK) or len(chunk) >= DOT_MAX_TERMS):
                part = do_dot(chunk)
                acc = part if acc is None else do_add(acc, part)
                chunk, chunk_total, chunk_slots = [], 0, 0
            chunk.append((a, const_opnd(c)))
            chunk_total += w
            chunk_slots += a.slot is not None
        if chunk:
            part = do_dot(chunk)
            acc = part if acc is None else do_add(acc, part)
        if c0:
            acc = const_opnd(c0) if acc is None else do_add(acc, const_opnd(c0))
        for t in plus:
            op = operand(t)
            acc = op if acc is None else do_add(acc, op)
        for t in minus:
            op = operand(t)
            acc = do_sub(acc if acc is not None else const_opnd(0), op)
        for t, sm in small:
            m = multiple(operand(t), abs(sm))
            if sm > 0:
                acc = m if acc is None else do_add(acc, m)
            else:
                acc = do_sub(acc if acc is not None else const_opnd(0), m)
        adopt(acc, nid)

    def emit(nid: int):
        k, a = g.kind[nid], g.args[nid]
        if k == "x":
            d = alloc()
            code.append((OP_MOV | d << 8, K_X << 29, 0, 0))
            bound[d] = X_BOUND
            slot_of[nid] = d
            stats["mul"] += 1
        elif k == "mul":
            x = operand(a[0])
            y = operand(a[1])
            adopt(do_mul(x, y), nid)
        elif k == "inv":
            x = operand(a[0])
            if reducible(x) and x.b > OUT_CAP:
                red(x)
            adopt(put(OP_INV, 0, x, None, _mul_bound(OUT_CAP, OUT_CAP)), nid)
            stats["mul"] += 262
            stats["inv"] += 1
        elif k == "lc":
            emit_lc(nid)
        else:
            raise ValueError(k)

    out = operand(root_node)
    if out.b > 4 * P:
        red(out)
    code.append((OP_OUT, out.word, 0, 0))
    if n_slots > max_slots:
        raise ValueError(f"program needs {n_slots} value slots, the kernel has {max_slots}")

    # ---- 5. serialise --------------------------------------------------------------------------------------
    # A table whose expression involves a tracked constant must have the form  constant * (challenge-independent table):
    # the values of the second factor are stored and the constant becomes the table's device-side scale.
    sym_memo: dict = {}

    def has_sym(e: Expr) -> bool:
        hit = sym_memo.get(e)
        if hit is None:
            hit = isinstance(e.args[0], Sym) if e.op == "const" else any(has_sym(a) for a in e.args if isinstance(a, Expr))
            sym_memo[e] = hit
        return hit

    def split_sym(e: Expr):
        """e == factor * base with `factor` the product of the tracked constants among e's multiplicative factors
        (None = 1) and `base` challenge-independent (None = 1)."""
        if not has_sym(e):
            return None, e
        if e.op == "const":
            return e.args[0], None
        if e.op in ("mul", "div"):
            (fa, ba), (fb, bb) = split_sym(e.args[0]), split_sym(e.args[1])
            if e.op == "div":
                fb = pow(fb, -1, P) if fb is not None else None
                bb = Expr("div", Expr("const", 1), bb) if bb is not None else None
            f = fa if fb is None else (fb if fa is None else fa * fb % P)
            b = ba if bb is None else (bb if ba is None else Expr("mul", ba, bb))
            return f, b
        if e.op == "neg":
            f, b = split_sym(e.args[0])
            return (f, Expr("neg", b)) if b is not None else (-f % P, None)
        raise NotImplementedError("a periodic table depends on the challenges other than through a constant factor")

    table_memo: dict = {}
    table_vals, table_scale = [], []
    for e in tables:
        scale, base = None, e
        if has_sym(e):
            f, base = split_sym(e)
            scale = const_id(f)
            if base is None:
                base = Expr("const", 1)
        T = lw.period(e)
        # with_tables=False (build-time code generation): only the shape of the tables is needed
        vals = lw.table_values(base, table_memo) if with_tables else [0] * T
        if len(vals) != T:
            vals = vals * (T // len(vals))
        table_vals.append(vals)
        table_scale.append(scale)
    del table_memo
    if log_n + log_blowup > 32:
        raise ValueError("tap offsets are stored as 32-bit row indices")
    words = [MAGIC, VERSION, len(code), len(consts), len(tables), max(n_slots, 1), log_n, log_blowup, len(tap_list), 0, 0, 0, 0, 0, 0, 0]
    off = 0
    for t, sc in zip(table_vals, table_scale):
        words += [(len(t).bit_length() - 1) | ((sc + 1) << 8 if sc is not None else 0), off]
        off += len(t)
    if len(tables) & 1:
        words += [0, 0]                      # keep the following areas 16-byte aligned
    for col, toff in tap_list:
        words += [col, toff]
    if len(tap_list) & 1:
        words += [0, 0]
    for w4 in code:
        words += [w & 0xFFFFFFFF for w in w4]
    head = struct.pack(f"<{len(words)}I", *words)
    if len(head) % 32:
        head += b"\0" * (32 - len(head) % 32)
    tfelts = [l for t in table_vals for v in t for l in _mont_limbs(v)]
    table_bytes = np.array(tfelts, dtype=np.uint64).tobytes() if tfelts else b""
    stats_obj = CompiledProgram(blob=b"", n_instr=len(code), n_consts=len(consts), n_tables=len(tables), n_slots=max(n_slots, 1),
                                n_mul=stats["mul"], n_addsub=stats["addsub"], n_trace_taps=stats["trace"], n_batch_inv=len(inv_nodes),
                                table_sizes=[len(t) for t in table_vals], n_red=stats["red"], n_dot=stats["dot"], n_inv=stats["inv"])
    if _tape is not None:
        ops, refs = _tape.compact(consts)
        return ProgramTemplate(head=head, const_refs=refs, tape_ops=ops,
                               n_inputs=(len(lw.challenges), len(lw.hints), len(lw.coeffs)), table_bytes=table_bytes, stats=stats_obj)
    assert not any(isinstance(v, Sym) for v in consts)
    felts = [l for v in consts for l in _mont_limbs(v)]
    stats_obj.blob = head + (np.array(felts, dtype=np.uint64).tobytes() if felts else b"") + table_bytes
    return stats_obj


def structure_hash(blob: bytes) -> int:
    """64-bit FNV-1a over the structural part of a program blob: counts, blowup, table periods, tap columns and
    the code words — everything except the trace length, the tap row offsets and the constant / table VALUES.
    ss_constraint_eval computes the same hash to pick a kernel specialised at build time (tools/gen_ce_kernels.py)."""
    w = struct.unpack_from("<16I", blob, 0)
    n_words, n_consts, n_tables, n_slots, n_taps = w[2], w[3], w[4], w[5], w[8]
    nt, ntap = n_tables + (n_tables & 1), n_taps + (n_taps & 1)
    tdesc = struct.unpack_from(f"<{2 * n_tables}I", blob, 64)
    taps = struct.unpack_from(f"<{2 * n_taps}I", blob, 64 + 8 * nt)
    code = struct.unpack_from(f"<{4 * n_words}I", blob, 64 + 8 * nt + 8 * ntap)
    stream = [n_words, n_consts, n_tables, n_slots, w[7], n_taps] + [tdesc[2 * t] for t in range(n_tables)] + [taps[2 * t] for t in range(n_taps)] + list(code)
    h = 0xCBF29CE484222325
    for v in stream:
        for sh in (0, 8, 16, 24):
            h = ((h ^ ((v >> sh) & 0xFF)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def tap_reach(blob: bytes, log_rows: int) -> dict:
    """{column: (most negative, most positive)} signed row offsets (in LDE rows) of the taps of a program blob — what a
    rank that evaluates a row range has to hold of each column beyond the range itself (multi-GPU halos)."""
    w = struct.unpack_from("<16I", blob, 0)
    n_tables, n_taps = w[4], w[8]
    nt = n_tables + (n_tables & 1)
    taps = struct.unpack_from(f"<{2 * n_taps}I", blob, 64 + 8 * nt)
    N = 1 << log_rows
    out: dict = {}
    for t in range(n_taps):
        col, off = taps[2 * t], taps[2 * t + 1]
        if off >= N // 2:
            off -= N
        lo, hi = out.get(col, (0, 0))
        out[col] = (min(lo, off), max(hi, off))
    return out
