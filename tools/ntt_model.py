#!/usr/bin/env python3
"""Executable model of the multi-pass tile NTT implemented in sandstorm_b200/csrc/ntt_fp252.cu.

Pure-Python big ints, tiny tile sizes; used to validate the pass plan, the tile geometry, the
inter-pass twiddle exponents and the fused LDE before the CUDA transcription is run on a GPU.
Run: python tools/ntt_model.py
"""
import random
import sys

P = 2**251 + 17 * 2**192 + 1


def root(n):
    return pow(3, (P - 1) // n, P)


def brev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def plan(log_n, log_tile):
    """Split log_n into passes of <= log_tile stages, as evenly as possible, largest first."""
    if log_n <= log_tile:
        return [log_n]
    k = -(-log_n // log_tile)
    base, extra = divmod(log_n, k)
    return [base + (1 if i < extra else 0) for i in range(k)]


def dif_pass(x, log_b, L, w_n, log_n):
    """One DIF pass on every block of size 2^log_b: L stages on the top bits (local twiddles only),
    then the inter-pass twiddle w_B^(lo * brev_L(m)).  Natural -> (partially) bit-reversed."""
    B, S = 1 << log_b, 1 << (log_b - L)
    w_b = pow(w_n, 1 << (log_n - log_b), P)          # w_B
    w_loc = pow(w_b, S, P)                            # w_{2^L}
    for blk in range(0, len(x), B):
        for lo in range(S):
            v = [x[blk + m * S + lo] for m in range(1 << L)]
            for beta in range(L - 1, -1, -1):                      # local span 2^beta
                span = 1 << beta
                for l0 in range(1 << L):
                    if l0 & span:
                        continue
                    a, b = v[l0], v[l0 + span]
                    tw = pow(w_loc, (l0 & (span - 1)) << (L - 1 - beta), P)
                    v[l0], v[l0 + span] = (a + b) % P, (a - b) * tw % P
            for m in range(1 << L):
                if S > 1:
                    v[m] = v[m] * pow(w_b, lo * brev(m, L), P) % P
                x[blk + m * S + lo] = v[m]


def dit_pass(x, log_b, L, w_n, log_n):
    """Mirror: pre-twiddle w_B^(lo*brev_L(m)), then L DIT stages across m.  (partially) bitrev -> natural."""
    B, S = 1 << log_b, 1 << (log_b - L)
    w_b = pow(w_n, 1 << (log_n - log_b), P)
    w_loc = pow(w_b, S, P)
    for blk in range(0, len(x), B):
        for lo in range(S):
            v = [x[blk + m * S + lo] for m in range(1 << L)]
            if S > 1:
                v = [v[m] * pow(w_b, lo * brev(m, L), P) % P for m in range(1 << L)]
            for beta in range(L):
                span = 1 << beta
                for l0 in range(1 << L):
                    if l0 & span:
                        continue
                    tw = pow(w_loc, (l0 & (span - 1)) << (L - 1 - beta), P)
                    a, t = v[l0], v[l0 + span] * tw % P
                    v[l0], v[l0 + span] = (a + t) % P, (a - t) % P
            for m in range(1 << L):
                x[blk + m * S + lo] = v[m]


def ntt_dif(x, log_n, log_tile, inverse=False):
    w = root(1 << log_n)
    if inverse:
        w = pow(w, -1, P)
    log_b = log_n
    for L in plan(log_n, log_tile):
        dif_pass(x, log_b, L, w, log_n)
        log_b -= L


def ntt_dit(x, log_n, log_tile, inverse=False):
    w = root(1 << log_n)
    if inverse:
        w = pow(w, -1, P)
    passes = plan(log_n, log_tile)
    log_b = 0
    for L in reversed(passes):
        log_b += L
        dit_pass(x, log_b, L, w, log_n)


def naive(x, offset=1, inverse=False):
    n = len(x)
    w = root(n)
    if inverse:
        w = pow(w, -1, P)
    out = [sum(c * pow(offset * pow(w, i, P), k, P) for k, c in enumerate(x)) % P for i in range(n)]
    if inverse:
        ninv = pow(n, -1, P)
        out = [v * ninv % P for v in out]
    return out


def lde(evals, log_n, log_blowup, log_tile):
    """Fused LDE as the kernels do it: inverse DIF (natural -> bitrev coeffs), scale by
    n^-1 g^k with k = brev(pos), bit-reversed zero-padding = stride-b placement, forward DIT."""
    n, b = 1 << log_n, 1 << log_blowup
    c = list(evals)
    ntt_dif(c, log_n, log_tile, inverse=True)
    ninv = pow(n, -1, P)
    c = [v * ninv * pow(3, brev(p, log_n), P) % P for p, v in enumerate(c)]
    big = [0] * (n * b)
    for p, v in enumerate(c):
        big[p * b] = v
    ntt_dit(big, log_n + log_blowup, log_tile)
    return big


def main():
    rnd = random.Random(7)
    for log_tile in (3, 4):
        for log_n in range(1, 11):
            x = [rnd.randrange(P) for _ in range(1 << log_n)]
            want = naive(x)
            a = list(x)
            ntt_dif(a, log_n, log_tile)
            assert [a[brev(i, log_n)] for i in range(1 << log_n)] == want, ("dif", log_tile, log_n)
            b = [x[brev(i, log_n)] for i in range(1 << log_n)]
            ntt_dit(b, log_n, log_tile)
            assert b == want, ("dit", log_tile, log_n)
        for log_n, log_b in ((3, 1), (5, 1), (6, 2), (7, 1)):
            x = [rnd.randrange(P) for _ in range(1 << log_n)]
            coeffs = naive(x, inverse=True)
            want = naive(coeffs + [0] * ((1 << (log_n + log_b)) - len(x)), offset=3)
            assert lde(x, log_n, log_b, log_tile) == want, ("lde", log_tile, log_n, log_b)
    print("ntt_model: all plans OK")


if __name__ == "__main__":
    sys.exit(main())
