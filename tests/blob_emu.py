"""Python emulator of the `ss_constraint_eval` device interpreter (program blob version 2).

It follows the DEVICE semantics, not field semantics: values are Montgomery residues kept in the lazy
domain, additions are raw, SUBK adds k * p, RED and the Montgomery reductions use the exact formulas of
sandstorm_b200/csrc/fp252.cuh — and it asserts the invariants the compiler's bound analysis promises
(every stored value < 2^256, no borrow, unreduced accumulators < 2^512).  The result is returned as a
canonical field element, so that it can be compared with the independent tree evaluator."""
import struct

import numpy as np

P = 2**251 + 17 * 2**192 + 1
R = 2**256
RINV = pow(R, -1, P)
PINV = pow(P, -1, R)

OP_NOP, OP_MOV, OP_ADD, OP_SUBK, OP_RED, OP_MUL, OP_DOT, OP_INV, OP_OUT = range(9)
K_SLOT, K_CONST, K_TAP, K_TABLE, K_X = range(5)


def mont_reduce(S: int) -> int:
    """(S + p*R - q*p) / R with q = (S mod R) * p^-1 mod R  (fp252.cuh mont_reduce on T = S + p * 2^256)."""
    T = S + P * R
    assert T < 1 << 512, "accumulator overflow"
    q = (T % R) * PINV % R
    r, rem = divmod(T - q * P, R)
    assert rem == 0 and 0 <= r < R, "reduction result does not fit 256 bits"
    return r


def red(v: int) -> int:
    q = v >> 251
    return v - (q - (1 if q else 0)) * P


def run_blob(blob, i, lde_int, log_N, stats=None):
    """lde_int[col][row]: canonical field elements.  Returns the canonical output of row i."""
    w = struct.unpack_from("<16I", blob, 0)
    assert w[0] == 0x50435353 and w[1] == 3
    n_words, n_consts, n_tables, n_slots, n_taps = w[2], w[3], w[4], w[5], w[8]
    nt = n_tables + (n_tables & 1)
    ntap = n_taps + (n_taps & 1)
    tdesc = struct.unpack_from(f"<{2 * n_tables}I", blob, 64)
    taps = struct.unpack_from(f"<{2 * n_taps}I", blob, 64 + 8 * nt)
    code = struct.unpack_from(f"<{4 * n_words}I", blob, 64 + 8 * nt + 8 * ntap)
    head = (64 + 8 * nt + 8 * ntap + 16 * n_words + 31) // 32 * 32
    felts = np.frombuffer(blob, dtype=np.uint64, offset=head).reshape(-1, 4)
    raw = lambda k: int(felts[k][0]) | int(felts[k][1]) << 64 | int(felts[k][2]) << 128 | int(felts[k][3]) << 192
    N = 1 << log_N
    wN = pow(3, (P - 1) // N, P)
    s = [None] * n_slots

    def fetch(word):
        kind, pay = word >> 29, word & 0x1FFFFFFF
        if kind == K_SLOT:
            assert s[pay] is not None, "read of an unwritten slot"
            return s[pay]
        if kind == K_CONST:
            assert pay < n_consts
            return raw(pay)
        if kind == K_TAP:
            assert pay < n_taps
            col, off = taps[2 * pay], taps[2 * pay + 1]
            return lde_int[col][(i + off) % N] * R % P
        if kind == K_TABLE:
            lp, scale = tdesc[2 * pay] & 0xFF, tdesc[2 * pay] >> 8
            v = raw(n_consts + tdesc[2 * pay + 1] + (i & ((1 << lp) - 1)))
            return v * raw(scale - 1) * RINV % P if scale else v          # the device multiplies scaled tables out once
        if kind == K_X:
            return 3 * pow(wN, i, P) * R % P
        raise AssertionError(kind)

    out = None
    pc = 0
    while pc < n_words:
        w0, A, B, _ = code[4 * pc:4 * pc + 4]
        op, d, n = w0 & 0xFF, (w0 >> 8) & 0xFF, w0 >> 16
        pc += 1
        if op == OP_MOV:
            v = fetch(A)
        elif op == OP_ADD:
            v = fetch(A) + fetch(B)
        elif op == OP_SUBK:
            a, b = fetch(A), fetch(B)
            assert n <= 31 and b <= n * P, "SUBK bias too small"
            v = a - b + n * P
        elif op == OP_RED:
            v = red(fetch(A))
            assert v < 1 << 252
        elif op == OP_MUL:
            v = mont_reduce(fetch(A) * fetch(B))
        elif op == OP_DOT:
            S = 0
            for t in range(n):
                word = code[4 * (pc + t // 2):4 * (pc + t // 2) + 4]
                S += fetch(word[2 * (t & 1)]) * fetch(word[2 * (t & 1) + 1])
            pc += (n + 1) // 2
            v = mont_reduce(S)
        elif op == OP_INV:
            a = fetch(A)
            assert a < 8 * P + 8
            v = pow(a * RINV % P, -1, P) * R % P          # any representative below 3p is fine for the emulator
        elif op == OP_OUT:
            v = fetch(A)
            assert v < 4 * P, "OUT operand too large for canon()"
            out = v * RINV % P
            continue
        else:
            raise AssertionError(op)
        assert 0 <= v < R, f"value out of range after op {op}"
        s[d] = v
        if stats is not None:
            stats[op] = stats.get(op, 0) + (n if op == OP_DOT else 1)
    return out
