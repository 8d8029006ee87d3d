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


# ---- row-sharded transforms over W ranks (plan for the multi-GPU NTT) -------------------------------
# Every rank keeps a full-size array but only the positions it OWNS are valid.  Two ownership maps:
#   contiguous : position p belongs to rank p // (N / W)                      (what the row-sharded consumers want)
#   chunk-cyclic: position p belongs to rank (p // chunk) % W                  (chunk = the tile width of the strided passes)
# A strided pass touches, per tile, the positions blk + m * S + lo for all m: with S a multiple of chunk * W the owner of a
# position depends on lo only, so every tile of every strided pass is complete on one rank under the chunk-cyclic map —
# however many strided passes follow each other.  The contiguous pass (whole blocks of 2^L positions) is complete on one
# rank under the contiguous map.  Hence ONE all-to-all on each side of the contiguous pass:
#   DIF: contiguous -> [all-to-all] -> chunk-cyclic, strided passes -> [all-to-all] -> contiguous, last pass
#   DIT: contiguous, first pass -> [all-to-all] -> chunk-cyclic, strided passes -> [all-to-all] -> contiguous
# (condition: the smallest stride, i.e. the size 2^L of the contiguous pass, is >= chunk * W — 256+ against 8 * 8 in practice.)
# Each all-to-all moves (W-1)/W of a rank's 1/W share.  The functions below run the model passes per rank on exactly
# the tiles that rank owns and exchange ownership explicitly; unowned positions hold None so that any access to data a
# rank does not have raises.

def owner_contig(p, n, W):
    return p // (n // W)


def owner_cyclic(p, chunk, W):
    return (p // chunk) % W


def exchange(arrays, n, W, new_owner):
    """all-to-all: afterwards rank r holds exactly the positions with new_owner(p) == r."""
    merged = [None] * n
    for a in arrays:
        for p, v in enumerate(a):
            if v is not None:
                assert merged[p] is None, "two ranks own the same position"
                merged[p] = v
    assert all(v is not None for v in merged)
    return [[merged[p] if new_owner(p) == r else None for p in range(n)] for r in range(W)]


def dif_pass_owned(x, log_b, L, w_n, log_n, owns):
    """dif_pass restricted to the tiles whose positions satisfy owns(position) (checked on the tile's first element)."""
    B, S = 1 << log_b, 1 << (log_b - L)
    w_b = pow(w_n, 1 << (log_n - log_b), P)
    w_loc = pow(w_b, S, P)
    for blk in range(0, len(x), B):
        for lo in range(S):
            if not owns(blk + lo):
                continue
            v = [x[blk + m * S + lo] for m in range(1 << L)]
            assert all(e is not None for e in v), "tile not complete on this rank"
            for beta in range(L - 1, -1, -1):
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


def dit_pass_owned(x, log_b, L, w_n, log_n, owns):
    B, S = 1 << log_b, 1 << (log_b - L)
    w_b = pow(w_n, 1 << (log_n - log_b), P)
    w_loc = pow(w_b, S, P)
    for blk in range(0, len(x), B):
        for lo in range(S):
            if not owns(blk + lo):
                continue
            v = [x[blk + m * S + lo] for m in range(1 << L)]
            assert all(e is not None for e in v), "tile not complete on this rank"
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


def sharded_dif(x, log_n, log_tile, W, chunk, inverse=False):
    """natural order, contiguous shards in -> bit-reversed order, contiguous shards out.  Returns the per-rank arrays."""
    n = 1 << log_n
    w = root(n)
    if inverse:
        w = pow(w, -1, P)
    passes = plan(log_n, log_tile)
    # the smallest stride of a strided pass is the size of the last (contiguous) pass: it must cover one chunk per rank
    assert len(passes) == 1 or (1 << passes[-1]) >= chunk * W, "last pass too small for this chunk-cyclic ownership"
    arrays = [[v if owner_contig(p, n, W) == r else None for p, v in enumerate(x)] for r in range(W)]
    log_b, state = log_n, "contig"
    for k, L in enumerate(passes):
        last = k == len(passes) - 1
        want = "contig" if last else "cyclic"
        if want != state:
            arrays = exchange(arrays, n, W, (lambda p: owner_contig(p, n, W)) if want == "contig" else (lambda p: owner_cyclic(p, chunk, W)))
            state = want
        for r in range(W):
            owns = (lambda p, r=r: owner_contig(p, n, W) == r) if state == "contig" else (lambda p, r=r: owner_cyclic(p, chunk, W) == r)
            dif_pass_owned(arrays[r], log_b, L, w, log_n, owns)
        log_b -= L
    return arrays


def sharded_dit(arrays, log_n, log_tile, W, chunk, inverse=False):
    """bit-reversed order, contiguous shards in -> natural order, contiguous shards out (in place on the rank arrays)."""
    n = 1 << log_n
    w = root(n)
    if inverse:
        w = pow(w, -1, P)
    passes = plan(log_n, log_tile)
    log_b, state = 0, "contig"
    for k, L in enumerate(reversed(passes)):
        log_b += L
        want = "contig" if k == 0 else "cyclic"
        if want != state:
            arrays = exchange(arrays, n, W, (lambda p: owner_contig(p, n, W)) if want == "contig" else (lambda p: owner_cyclic(p, chunk, W)))
            state = want
        for r in range(W):
            owns = (lambda p, r=r: owner_contig(p, n, W) == r) if state == "contig" else (lambda p, r=r: owner_cyclic(p, chunk, W) == r)
            dit_pass_owned(arrays[r], log_b, L, w, log_n, owns)
    if state != "contig":
        arrays = exchange(arrays, n, W, lambda p: owner_contig(p, n, W))
    return arrays


def gather(arrays):
    n = len(arrays[0])
    out = [None] * n
    for a in arrays:
        for p, v in enumerate(a):
            if v is not None:
                out[p] = v
    return out


def check_sharded(rnd):
    for log_tile, log_n, W, chunk in ((3, 7, 2, 2), (3, 8, 4, 1), (4, 9, 4, 2), (3, 9, 8, 1), (4, 10, 2, 4)):
        n = 1 << log_n
        x = [rnd.randrange(P) for _ in range(n)]
        want = naive(x)
        got = gather(sharded_dif(x, log_n, log_tile, W, chunk))
        assert [got[brev(i, log_n)] for i in range(n)] == want, ("sharded dif", log_tile, log_n, W)
        xb = [x[brev(i, log_n)] for i in range(n)]
        arrays = [[v if owner_contig(p, n, W) == r else None for p, v in enumerate(xb)] for r in range(W)]
        assert gather(sharded_dit(arrays, log_n, log_tile, W, chunk)) == want, ("sharded dit", log_tile, log_n, W)


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
    check_sharded(rnd)
    print("ntt_model: all plans OK (single-rank and row-sharded)")


if __name__ == "__main__":
    sys.exit(main())
