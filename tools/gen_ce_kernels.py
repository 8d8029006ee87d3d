#!/usr/bin/env python3
"""Ahead-of-time specialisation of the constraint-program interpreter.

`ss_constraint_eval` executes a program blob (sandstorm_b200/air/program.py).  For the programs whose
STRUCTURE is known at build time — the composition constraint and the DEEP quotient of each reference
layout (layouts/src/{plain,recursive,starknet}/air.rs) — this script translates the code words 1:1 into
straight-line CUDA: slots become registers, operand kinds and indices become literals, Montgomery
multiplication stays one shared non-inlined routine (so that the instruction stream is short), and ptxas
schedules the loads.  Constants, tables, tap offsets and matrix pointers still come from the blob /
arguments at run time, so one generated kernel serves every trace length and every challenge draw.

The kernels are registered under a 64-bit FNV-1a hash of the structural part of the blob; at run time
ss_constraint_eval computes the same hash and launches the specialised kernel when one exists, else the
interpreter — the result is bit-identical either way (same operations in the same order).

Usage: python tools/gen_ce_kernels.py OUT.cuh   (invoked by sandstorm_b200/csrc/Makefile)"""
from __future__ import annotations

import os
import random
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from sandstorm_b200.air import compile_program, compile_template  # noqa: E402
from sandstorm_b200.air.deep import DEEP_FILTER_MIN_TAPS, deep_expr_filtered, deep_expr_shifted, deep_filter_columns, deep_terms  # noqa: E402
from sandstorm_b200.air.expr import P  # noqa: E402
from sandstorm_b200.air.layouts import load_layout  # noqa: E402
from sandstorm_b200.air.program import structure_hash  # noqa: E402

OP_NOP, OP_MOV, OP_ADD, OP_SUBK, OP_RED, OP_MUL, OP_DOT, OP_INV, OP_OUT = range(9)
K_SLOT, K_CONST, K_TAP, K_TABLE, K_X = range(5)


def decode(blob: bytes):
    w = struct.unpack_from("<16I", blob, 0)
    n_words, n_consts, n_tables, n_slots, n_taps = w[2], w[3], w[4], w[5], w[8]
    nt, ntap = n_tables + (n_tables & 1), n_taps + (n_taps & 1)
    tdesc = struct.unpack_from(f"<{2 * n_tables}I", blob, 64)
    taps = struct.unpack_from(f"<{2 * n_taps}I", blob, 64 + 8 * nt)
    code = struct.unpack_from(f"<{4 * n_words}I", blob, 64 + 8 * nt + 8 * ntap)
    return dict(n_words=n_words, n_consts=n_consts, n_tables=n_tables, n_slots=n_slots, n_taps=n_taps, tdesc=tdesc, taps=taps, code=code)


def operand(word: int, d: dict) -> str:
    kind, pay = word >> 29, word & 0x1FFFFFFF
    if kind == K_SLOT:
        return f"s{pay}"
    if kind == K_CONST:
        return f"ldg_fp(A.consts + {pay})"
    if kind == K_TAP:
        return f"ldg_fp(A.cols + {d['taps'][2 * pay]}ull * A.stride + ((i + G.tap_off[{pay}]) & mask))"
    if kind == K_TABLE:
        return f"ldg_fp(A.tables + G.tab_off[{pay}] + (i & {(1 << (d['tdesc'][2 * pay] & 0xFF)) - 1}ull))"
    if kind == K_X:
        return "fetch_x(A, i)"
    raise ValueError(kind)


def emit_kernel(name: str, blob: bytes, minb: int) -> tuple[str, int]:
    d = decode(blob)
    code = d["code"]
    out = [f"// {name}: {d['n_words']} code words, {d['n_slots']} slots, {d['n_taps']} taps, {d['n_tables']} tables",
           f"__global__ void __launch_bounds__(128, {minb}) ce_gen_{name}(const EvalArgs A, const __grid_constant__ GenArgs G) {{",
           "    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;",
           "    if (t >= A.row_count) return;",
           "    const unsigned long long i = A.row_begin + (t << A.log_step), mask = (1ull << A.log_N) - 1;",
           "    Fp " + ", ".join(f"s{k}" for k in range(d["n_slots"])) + ";"]
    pc = 0
    while pc < d["n_words"]:
        w0, a, b, _ = code[4 * pc:4 * pc + 4]
        op, dst, n = w0 & 0xFF, (w0 >> 8) & 0xFF, w0 >> 16
        pc += 1
        if op == OP_MOV:
            out.append(f"    s{dst} = {operand(a, d)};")
        elif op == OP_ADD:
            out.append(f"    s{dst} = fp::add_raw({operand(a, d)}, {operand(b, d)});")
        elif op == OP_SUBK:
            out.append(f"    s{dst} = fp::sub_kp({operand(a, d)}, {operand(b, d)}, {n}u);")
        elif op == OP_RED:
            out.append(f"    s{dst} = fp::red(s{dst});")
        elif op == OP_MUL:
            out.append(f"    s{dst} = mul_ni({operand(a, d)}, {operand(b, d)});")
        elif op == OP_DOT:
            out.append("    {")
            for k in range(n):
                word = code[4 * (pc + k // 2):4 * (pc + k // 2) + 4]
                x, y = operand(word[2 * (k & 1)], d), operand(word[2 * (k & 1) + 1], d)
                out.append(f"        Wide acc = wide_p_ni({x}, {y});" if k == 0 else f"        wide_add(acc, wide_ni({x}, {y}));")
            pc += (n + 1) // 2
            out.append(f"        s{dst} = reduce_ni(acc);")
            out.append("    }")
        elif op == OP_INV:
            out.append(f"    s{dst} = ec::inv_chain({operand(a, d)});")
        elif op == OP_OUT:
            out.append(f"    store_out(A.out + (i >> A.log_step), fp::canon({operand(a, d)}));")
        elif op != OP_NOP:
            raise ValueError(op)
    out.append("}")
    return "\n".join(out), structure_hash(blob)


def programs():
    """(name, blob, resident CTAs per SM) for every program specialised at build time."""
    rnd = random.Random(0xB200)
    out = []
    for layout, log_n in (("starknet", 18), ("recursive", 14), ("plain", 10)):
        L = load_layout(layout)
        C, n = L.num_columns, 1 << log_n
        ce = 2
        w_col = C + ce                                         # HotPathProver column map: trace | composition | w | u | v
        comp = compile_template(L.composition(n, inv_x_minus_one_col=w_col), log_n, 1, L.n_challenges(), L.n_hints(), 1, with_tables=False)
        comp = comp.patch([rnd.randrange(P) for _ in range(L.n_challenges())], [rnd.randrange(P) for _ in range(L.n_hints())], [rnd.randrange(P)])
        out.append((f"{layout}_composition", comp.blob, 3))
        g = pow(3, (P - 1) // n, P)
        tt, ct = deep_terms(L.taps(), [rnd.randrange(P) for _ in L.taps()], [rnd.randrange(P) for _ in range(ce)], C, rnd.randrange(P), P)
        deep = compile_program(deep_expr_shifted(tt, ct, C + ce + 1, C + ce + 2, g, P), log_n, 1, with_tables=False)
        out.append((f"{layout}_deep", deep.blob, 4))
        # the single-GPU prover's form: long pole sums read from precomputed columns (HotPathProver.__init__ column map)
        taps = L.taps()
        heavy = deep_filter_columns(taps, DEEP_FILTER_MIN_TAPS)
        if heavy or len({off for _, off in taps}) >= DEEP_FILTER_MIN_TAPS:
            value_col = C + ce + 3
            fcols = {col: value_col + 1 + j for j, col in enumerate(heavy)}
            t = compile_template(deep_expr_filtered(taps, ce, C, C + ce + 1, C + ce + 2, g, P, fcols, value_col), log_n, 1, 1, len(taps) + ce, 1,
                                 with_tables=False)
            deepf = t.patch([rnd.randrange(P)], [rnd.randrange(P) for _ in range(len(taps) + ce)], [0])
            out.append((f"{layout}_deepf", deepf.blob, 4))
    return out


def main():
    dst = sys.argv[1]
    parts = ["// GENERATED by tools/gen_ce_kernels.py — do not edit; included by constraint_eval.cu", ""]
    reg = []
    variants = [int(v) for v in os.environ.get("SS_GEN_MINB_VARIANTS", "").split(",") if v]
    for name, blob, minb in programs():
        for mb in (variants if variants and name.startswith("starknet") else [minb]):
            src, h = emit_kernel(f"{name}_mb{mb}", blob, mb)
            parts.append(src)
            parts.append("")
            reg.append((f"{name}_mb{mb}", h, mb, mb == minb or (variants and mb == variants[0] and minb not in variants)))
    parts.append("struct GenEntry { unsigned long long hash; void (*kernel)(const EvalArgs, const GenArgs); const char *name; int minb; bool dflt; };")
    parts.append("const GenEntry GEN_KERNELS[] = {")
    for name, h, mb, dflt in reg:
        parts.append(f"    {{0x{h:016x}ull, ce_gen_{name}, \"{name}\", {mb}, {'true' if dflt else 'false'}}},")
    parts.append("};")
    tmp = dst + ".tmp"
    with open(tmp, "w") as f:
        f.write("\n".join(parts) + "\n")
    os.replace(tmp, dst)


if __name__ == "__main__":
    main()
