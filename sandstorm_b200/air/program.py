"""Compiler: composition-constraint Expr DAG -> straight-line program for `ss_constraint_eval`.

What the reference does at run time with an expression interpreter over whole vectors (ministark's
evaluator) is split here between host and device:

  host   : substitute challenges / hints / composition coefficient, fold constants, hash-cons (CSE),
           classify every node by its PERIOD in the LDE row index i (x_i = 3 * w_N^i):
             period 1        -> constant
             period T <= 2^17 -> lookup table of T entries (zerofiers X^(n/k) - c and their inverses,
                                periodic columns P(X^(n/interval)), products of those)
             full            -> depends on trace cells or on X itself -> device instruction
           full-period denominators (the X - g^e boundary terms) are inverted together per row with
           one batched inversion.
  device : one thread per LDE row runs the instruction list over a small slot file (liveness-allocated).

Program blob layout (little-endian u32 words unless noted):
  [0] magic 'SSCP'  [1] version  [2] n_instr  [3] n_consts  [4] n_tables  [5] n_slots  [6] log_n (trace)
  [7] log_blowup    then n_tables x (log_period, offset in elements), padded to an even count
  then n_instr x 4 words (op | dst << 8, a, b, imm)    then consts (32 B each)   then table data (32 B each)
"""
from __future__ import annotations

import math
import struct
from dataclasses import dataclass, field

import numpy as np

from .expr import Expr, P

R = 2**256
MAGIC = 0x50435353          # 'SSCP'
VERSION = 1
MAX_TABLE_LOG = 17
GENERATOR = 3

OP_NOP, OP_CONST, OP_TRACE, OP_TABLE, OP_X, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_INV, OP_BATCHINV, OP_OUT, OP_MULC, OP_ADDC = range(14)
FULL = 0    # period marker for row-dependent nodes


def _mont_limbs(v: int) -> list[int]:
    m = v % P * R % P
    return [(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


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


def compile_program(expr: Expr, log_n: int, log_blowup: int, challenges=(), hints=(), composition_coeffs=(0,),
                    max_slots: int = 256) -> CompiledProgram:
    """expr: the composition constraint (or any Expr).  challenges / hints / composition_coeffs: canonical ints."""
    lw = _Lower(log_n, log_blowup, list(challenges), list(hints), list(composition_coeffs))
    root = lw.norm(expr)
    N = lw.N

    consts: list[int] = []
    const_ix: dict[int, int] = {}
    tables: list[list[int]] = []
    table_ix: dict[Expr, int] = {}

    def const_id(v: int) -> int:
        v %= P
        if v not in const_ix:
            const_ix[v] = len(consts)
            consts.append(v)
        return const_ix[v]

    def table_id(e: Expr) -> int:
        if e not in table_ix:
            T = lw.period(e)
            table_ix[e] = len(tables)
            tables.append([lw.eval_at(e, j) for j in range(T)])
            lw.memo_eval.clear()
        return table_ix[e]

    # ---- lower the full-period part into a DAG of device nodes (hash-consed tuples) -----------------------
    dev_memo: dict[Expr, tuple] = {}
    inv_nodes: list[tuple] = []          # ('inv', denominator_node) for full-period denominators

    def lower(e: Expr) -> tuple:
        hit = dev_memo.get(e)
        if hit is not None:
            return hit
        p = lw.period(e)
        if p == 1:
            node = ("const", const_id(lw.eval_at(e, 0)))
        elif p != FULL:
            node = ("table", table_id(e))
        elif e.op == "x":
            node = ("x",)
        elif e.op == "trace":
            node = ("trace", e.args[0], e.args[1] * (1 << log_blowup))
        elif e.op == "pow":                                      # X^k with a long period: square-and-multiply on X
            k, acc, sq = e.args[1], None, ("x",)
            while k:
                if k & 1:
                    acc = sq if acc is None else ("mul", acc, sq)
                k >>= 1
                if k:
                    sq = ("mul", sq, sq)
            node = acc
        elif e.op == "periodic":
            raise ValueError("periodic column with a period above the table limit")
        elif e.op == "div":
            num, den = e.args
            if lw.period(den) != FULL:
                node = ("mul", lower(num), lower(Expr("div", Expr("const", 1), den)))
            else:
                inv = ("inv", lower(den))
                if inv not in inv_nodes:
                    inv_nodes.append(inv)
                node = ("mul", lower(num), inv)
        elif e.op == "neg":
            node = ("neg", lower(e.args[0]))
        else:
            node = (e.op, lower(e.args[0]), lower(e.args[1]))
        dev_memo[e] = node
        return node

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    root_node = lower(root)

    # ---- schedule: denominators first (pinned contiguous slots), one BATCHINV, then the rest -----------
    dep_memo: dict = {}

    def depends_on_inv(node):
        if node in dep_memo:
            return dep_memo[node]
        r = node[0] == "inv" or any(isinstance(c, tuple) and depends_on_inv(c) for c in node[1:])
        dep_memo[node] = r
        return r

    batch = [iv for iv in inv_nodes if not depends_on_inv(iv[1])][:192]
    batch_set = set(batch)
    code: list[tuple] = []
    slot_of: dict[tuple, int] = {}
    free: list[int] = []
    n_slots = 0
    stats = {"mul": 0, "addsub": 0, "trace": 0}

    def alloc() -> int:
        nonlocal n_slots
        if free:
            return free.pop()
        n_slots += 1
        return n_slots - 1

    # use counts for liveness
    uses: dict[tuple, int] = {}

    def count(node):
        for c in node[1:]:
            if isinstance(c, tuple):
                uses[c] = uses.get(c, 0) + 1
                if uses[c] == 1:
                    count(c)

    roots = [iv[1] for iv in batch] + [root_node]
    for r in roots:
        uses[r] = uses.get(r, 0) + 1
        if uses[r] == 1:
            count(r)
    for iv in batch:                       # the inverse itself is consumed by its users
        pass

    pinned: set[int] = set()

    def release(node):
        uses[node] -= 1
        if node[0] in ("const", "table", "x", "trace"):
            if leaf_slot:
                free.append(leaf_slot.pop())        # the slot this use materialised
            return
        if uses[node] == 0 and node in slot_of:
            s = slot_of[node]
            if s not in pinned:
                free.append(s)

    LEAVES = ("const", "table", "x", "trace")

    def emit(node) -> int:
        # leaves are rematerialised at every use (a 32-byte L1/L2 hit is cheaper than a live slot):
        # only interior nodes are kept alive across uses
        if node in slot_of and node[0] not in LEAVES:
            return slot_of[node]
        kind = node[0]
        if kind == "inv" and node in batch_set:
            return slot_of[node]                                  # produced by BATCHINV
        if kind == "const":
            d = alloc(); code.append((OP_CONST, d, node[1], 0, 0))
        elif kind == "table":
            d = alloc(); code.append((OP_TABLE, d, node[1], 0, 0))
        elif kind == "x":
            d = alloc(); code.append((OP_X, d, 0, 0, 0))
        elif kind == "trace":
            d = alloc(); code.append((OP_TRACE, d, node[1], 0, node[2])); stats["trace"] += 1
        elif kind == "neg":
            a = emit(node[1]); release(node[1]); d = alloc(); code.append((OP_NEG, d, a, 0, 0)); stats["addsub"] += 1
        elif kind == "inv":
            a = emit(node[1]); release(node[1]); d = alloc(); code.append((OP_INV, d, a, 0, 0)); stats["mul"] += 262
        else:
            l, r = node[1], node[2]
            # constant operand forms save a slot and an instruction
            if kind == "mul" and (l[0] == "const" or r[0] == "const") and not (l[0] == "const" and r[0] == "const"):
                c, o = (l, r) if l[0] == "const" else (r, l)
                a = emit(o); release(o); uses[c] -= 1
                d = alloc(); code.append((OP_MULC, d, a, c[1], 0)); stats["mul"] += 1
            elif kind == "add" and (l[0] == "const" or r[0] == "const") and not (l[0] == "const" and r[0] == "const"):
                c, o = (l, r) if l[0] == "const" else (r, l)
                a = emit(o); release(o); uses[c] -= 1
                d = alloc(); code.append((OP_ADDC, d, a, c[1], 0)); stats["addsub"] += 1
            else:
                # evaluate the deeper operand first (shorter live ranges)
                a = emit(l); b = emit(r)
                release(l); release(r)
                d = alloc()
                code.append(({"add": OP_ADD, "sub": OP_SUB, "mul": OP_MUL}[kind], d, a, b, 0))
                stats["mul" if kind == "mul" else "addsub"] += 1
        if kind in LEAVES:
            leaf_slot.append(d)
        else:
            slot_of[node] = d
        return d

    leaf_slot: list[int] = []

    def emit_operand(node) -> int:
        return emit(node)

    if batch:
        base = n_slots
        for k, iv in enumerate(batch):     # reserve contiguous pinned slots [base, base + len)
            n_slots += 1
            pinned.add(base + k)
        for k, iv in enumerate(batch):
            den = iv[1]
            s = emit(den)
            # move into the pinned slot (ADDC 0); the denominator's own slot is recycled at once — if it
            # is also used as a plain factor later it is recomputed (one subtraction)
            code.append((OP_ADDC, base + k, s, const_id(0), 0))
            if den[0] in LEAVES:
                release(den)
            else:
                uses[den] -= 1
                if s not in pinned:
                    free.append(s)
                slot_of.pop(den, None)
            slot_of[iv] = base + k
        code.append((OP_BATCHINV, 0, base, len(batch), 0))
        stats["mul"] += 262 + 3 * (len(batch) - 1)
    out_slot = emit(root_node)
    code.append((OP_OUT, 0, out_slot, 0, 0))
    if n_slots > max_slots:
        raise ValueError(f"program needs {n_slots} value slots, the kernel has {max_slots}")

    # ---- serialise ------------------------------------------------------------------------------------
    words = [MAGIC, VERSION, len(code), len(consts), len(tables), n_slots, log_n, log_blowup]
    off = 0
    for t in tables:
        words += [len(t).bit_length() - 1, off]
        off += len(t)
    if len(tables) & 1:
        words += [0, 0]                      # keep the instruction area 16-byte aligned
    for op, d, a, b, imm in code:
        words += [op | (d << 8), a & 0xFFFFFFFF, b & 0xFFFFFFFF, imm & 0xFFFFFFFF]
    head = struct.pack(f"<{len(words)}I", *words)
    if len(head) % 32:
        head += b"\0" * (32 - len(head) % 32)
    felts = [l for v in consts for l in _mont_limbs(v)] + [l for t in tables for v in t for l in _mont_limbs(v)]
    body = np.array(felts, dtype=np.uint64).tobytes() if felts else b""
    return CompiledProgram(blob=head + body, n_instr=len(code), n_consts=len(consts), n_tables=len(tables), n_slots=n_slots,
                           n_mul=stats["mul"], n_addsub=stats["addsub"], n_trace_taps=stats["trace"], n_batch_inv=len(batch),
                           table_sizes=[len(t) for t in tables])
