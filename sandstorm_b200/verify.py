"""Host-side verification of Merkle openings — `MerkleTree::verify` / `MatrixMerkleTree::verify_rows` of the reference's tree
variants (crypto/src/merkle/mod.rs:125-165 Friendly, :306-346 LeafVariant; level rules crypto/src/merkle/mixed.rs:110-155).
The reference verifies on the CPU too (sandstorm verify, cli/src/main.rs:168-178): a few hundred hashes per proof.

Node numbering and depth follow the tree builder (csrc/merkle.cu): the node above leaves 2i, 2i+1 has depth log2(n) - 1;
a Friendly tree hashes levels of depth < n_friendly with Pedersen (children that are byte digests are first read as
big-endian integers, mixed.rs:148-155) and the levels below with masked Blake2s.  Paths are in storage form (what
`ss_merkle_open` returns): byte digests, or the Montgomery limbs of the felt for Pedersen levels."""
from __future__ import annotations

import numpy as np

from . import _lib
from . import hostcrypto as hc

P, R = hc.P, hc.R
_RINV = pow(R, -1, P)


def _limbs_to_int(b: bytes) -> int:
    return int.from_bytes(b, "little")


def _felt_from_storage(b: bytes) -> int:
    """32 bytes of Montgomery limbs (little-endian) -> canonical int."""
    return _limbs_to_int(b) * _RINV % P


def _felt_to_storage(v: int) -> bytes:
    return (v % P * R % P).to_bytes(32, "little")


def _byte_hash(kind: int):
    if kind == _lib.TREE_KECCAK:
        return hc.keccak256
    if kind == _lib.TREE_KECCAK_M20:
        return lambda d: hc.mask_keccak20(hc.keccak256(d))
    if kind in (_lib.TREE_BLAKE2S_M20, _lib.TREE_FRIENDLY):
        return lambda d: hc.mask_blake20(hc.blake2s(d))
    if kind == _lib.TREE_SHA256:
        return hc.sha256
    raise ValueError(kind)


def _be32_of_storage(b: bytes) -> bytes:
    """hash_elements encoding: big-endian bytes of the Montgomery limbs (crypto/src/utils.rs:15-17)."""
    return _limbs_to_int(b).to_bytes(32, "big")


def merkle_root_from_opening(kind: int, index: int, row: np.ndarray, path: np.ndarray, n_friendly: int = 22) -> bytes:
    """row: uint64[n_cols, 4] (the opened matrix row, Montgomery limbs); path: uint8[depth, 32], leaf level first.
    Returns the root as `Digest::as_bytes` (what ss_merkle_root returns)."""
    n_cols, height = row.shape[0], path.shape[0]
    H = _byte_hash(kind)
    friendly = kind == _lib.TREE_FRIENDLY
    sib = [bytes(path[k]) for k in range(height)]
    if n_cols == 1:
        # raw leaves (mod.rs:113-116, 292-295); first level = hash_elements of the pair (mod.rs:426-428)
        me = row[0].astype("<u8").tobytes()
        pair = (me, sib[0]) if index & 1 == 0 else (sib[0], me)
        if friendly:
            cur, algebraic = _felt_to_storage(hc.pedersen_hash_elements([_felt_from_storage(p) for p in pair])), True
        else:
            cur, algebraic = H(_be32_of_storage(pair[0]) + _be32_of_storage(pair[1])), False
        start = 1
    else:
        cur, algebraic, start = H(b"".join(_be32_of_storage(row[j].astype("<u8").tobytes()) for j in range(n_cols))), False, 0
    for k in range(start, height):
        depth = height - 1 - k                                   # depth of the parent built at this step
        left, right = (cur, sib[k]) if (index >> k) & 1 == 0 else (sib[k], cur)
        high = friendly and (n_cols == 1 or depth < n_friendly)
        if not high:
            cur, algebraic = H(left + right), False
        else:
            if algebraic:
                a, b = _felt_from_storage(left), _felt_from_storage(right)
            else:                                                # boundary: digest bytes -> big-endian integer -> felt
                a, b = int.from_bytes(left, "big") % P, int.from_bytes(right, "big") % P
            cur, algebraic = _felt_to_storage(hc.pedersen_hash(a, b)), True
    return _felt_from_storage(cur).to_bytes(32, "big") if algebraic else cur


# ---- conventions pinned by the reference's own proof artefacts (tests/test_reference_proof.py) ---------------------------------
# * every tree is built over rows in BIT-REVERSED order of the evaluation domain: leaf p commits the row at
#   x = 3 * w_N^brev(p); a query position p is such a leaf index (the same p for the three trace trees and for FRI);
# * a FRI layer commits rows of `fold` CONSECUTIVE entries of its bit-reversed evaluation vector, i.e. leaf r holds
#   f(x * w_F^brev(j)), j < F, for x = offset * w^brev(r): the F evaluations that fold together;
# * a fold is  sum_m alpha^m x^-m sum_k f(x w_F^k) w_F^(-m k)  — no 1/F factor (StarkWare's convention) — and lands at
#   entry r of the next layer (domain offset^F);
# * the remainder is sent as the coefficients of f(offset_last * X).
def brev(v: int, bits: int) -> int:
    return int(f"{v:0{bits}b}"[::-1], 2) if bits else 0


def root_from_leaf(kind: int, index: int, leaf, sibling, path, n_friendly: int = 22, unhashed: bool = False) -> bytes:
    """Recomputes a root from a serialized MerkleProof (proof.py): leaf / sibling are digests (bytes), or canonical felts
    (ints) for the raw single-column variant; path entries are bytes or, for Pedersen levels, canonical felts."""
    H = _byte_hash(kind)
    friendly = kind == _lib.TREE_FRIENDLY
    height = len(path) + 1
    if unhashed:
        pair = (leaf, sibling) if index & 1 == 0 else (sibling, leaf)
        if friendly:
            cur, algebraic = hc.pedersen_hash_elements(pair), True
        else:
            cur, algebraic = H(hc.felt_bytes(pair[0]) + hc.felt_bytes(pair[1])), False
    else:
        left, right = (leaf, sibling) if index & 1 == 0 else (sibling, leaf)
        high = friendly and height - 1 < n_friendly
        if high:
            cur, algebraic = hc.pedersen_hash(int.from_bytes(left, "big") % P, int.from_bytes(right, "big") % P), True
        else:
            cur, algebraic = H(left + right), False
    for k, sib in enumerate(path, start=1):
        depth = height - 1 - k
        left, right = (cur, sib) if (index >> k) & 1 == 0 else (sib, cur)
        high = friendly and (unhashed or depth < n_friendly)
        if not high:
            cur, algebraic = H(left + right), False
        else:
            a, b = (left, right) if algebraic else (int.from_bytes(left, "big") % P, int.from_bytes(right, "big") % P)
            cur, algebraic = hc.pedersen_hash(a, b), True
    return cur.to_bytes(32, "big") if algebraic else cur


def row_digest(kind: int, values) -> bytes:
    """hash_row (crypto/src/merkle/utils.rs:9-17): H over the BE32 Montgomery encodings of the row's elements."""
    return _byte_hash(kind)(b"".join(hc.felt_bytes(v) for v in values))


def fri_fold_row(values, r: int, log_domain: int, offset: int, alpha: int, log_fold: int = 3) -> int:
    """Folds one opened FRI row (leaf r of a layer over a domain of 2^log_domain points with the given offset)."""
    F = 1 << log_fold
    w = pow(3, (P - 1) >> log_domain, P)
    wF_inv = pow(3, -((P - 1) >> log_fold), P)
    x_inv = pow(offset * pow(w, brev(r, log_domain - log_fold), P) % P, -1, P)
    f = [values[brev(k, log_fold)] for k in range(F)]               # f(x * w_F^k)
    acc, am, xm = 0, 1, 1
    for m in range(F):
        acc = (acc + am * xm % P * sum(f[k] * pow(wF_inv, m * k, P) for k in range(F))) % P
        am, xm = am * alpha % P, xm * x_inv % P
    return acc


def remainder_at(coeffs, y: int, offset: int) -> int:
    t, acc = y * pow(offset, -1, P) % P, 0
    for c in reversed(coeffs):
        acc = (acc * t + c) % P
    return acc


# ---- the verifier: `Stark::verify` (ministark, not vendored) restated on the conventions above -----------------------------
class VerificationError(Exception):
    pass


def _as_bytes(d) -> bytes:
    """Digest::as_bytes: byte digests verbatim, a Pedersen felt as its big-endian canonical integer."""
    return d.to_bytes(32, "big") if isinstance(d, int) else bytes(d)


def _evaluate_at(e, z, tap_values, log_n, challenges, hints, coeffs):
    """the composition constraint at an out-of-domain point, Trace(col, off) := the claimed T_col(z g^off)."""
    import sys

    n = 1 << log_n
    memo = {}

    def go(e):
        if e in memo:
            return memo[e]
        op = e.op
        if op == "x": v = z
        elif op == "const": v = e.args[0]
        elif op == "trace": v = tap_values[(e.args[0], e.args[1])]
        elif op == "challenge": v = challenges[e.args[0]]
        elif op == "hint": v = hints[e.args[0]]
        elif op == "composition_coeff": v = coeffs[e.args[0]]
        elif op == "periodic":
            cs, interval = e.args
            y, v = pow(z, n // interval, P), 0
            for c in reversed(cs):
                v = (v * y + c) % P
        elif op == "pow": v = pow(go(e.args[0]), e.args[1], P)
        elif op == "neg": v = -go(e.args[0]) % P
        else:
            a, b = go(e.args[0]), go(e.args[1])
            v = (a + b if op == "add" else a - b if op == "sub" else a * b if op == "mul" else a * pow(b, -1, P)) % P
        memo[e] = v
        return v

    sys.setrecursionlimit(max(sys.getrecursionlimit(), 100000))
    return go(e)


def verify_proof(proof, layout, coin, gen_hints, tree_kind: int, n_friendly: int = 22, ce: int = 2) -> None:
    """Checks a Proof (sandstorm_b200/proof.py) of `layout` against a freshly seeded public coin: transcript, out-of-domain
    consistency of the composition polynomial, every Merkle opening, the DEEP quotient at every query and the FRI folds down
    to the remainder.  gen_hints: callable(challenges) -> hints (AirConfig::gen_hints over the public input).
    Raises VerificationError."""
    def need(cond, msg):
        if not cond:
            raise VerificationError(msg)

    n, b = proof.trace_len, proof.lde_blowup_factor
    log_n, log_b = n.bit_length() - 1, b.bit_length() - 1
    N, log_N = n * b, log_n + log_b
    F = proof.fri_folding_factor
    log_f = F.bit_length() - 1
    nb, ne = layout.num_base_columns, layout.num_extension_columns
    # transcript
    coin.reseed_with_digest(_as_bytes(proof.base_root))
    challenges = [coin.draw() for _ in range(layout.n_challenges())]
    hints = list(gen_hints(challenges))
    if proof.ext_root is not None:
        coin.reseed_with_digest(_as_bytes(proof.ext_root))
    comp_coeffs = [coin.draw()]
    coin.reseed_with_digest(_as_bytes(proof.comp_root))
    z = coin.draw()
    taps = layout.taps()
    need(len(proof.ood_trace) == len(taps) and len(proof.ood_comp) == ce, "wrong number of out-of-domain values")
    # out-of-domain consistency: C(z) = sum_j z^j comp_j(z^ce)
    lhs = _evaluate_at(layout.composition(n), z, dict(zip(taps, proof.ood_trace)), log_n, challenges, hints, comp_coeffs)
    need(lhs == sum(pow(z, j, P) * v for j, v in enumerate(proof.ood_comp)) % P, "out-of-domain evaluations are inconsistent with the AIR")
    coin.reseed_with_field_elements(proof.ood_trace)
    coin.reseed_with_field_elements(proof.ood_comp)
    deep_alpha = coin.draw()
    alphas = []
    for layer in proof.fri_layers:
        coin.reseed_with_digest(_as_bytes(layer.commitment))
        alphas.append(coin.draw())
    coin.reseed_with_field_element_vector(proof.remainder_coeffs)
    if proof.grinding_factor:
        need(coin.verify_proof_of_work(proof.grinding_factor, proof.pow_nonce), "proof of work")
        coin.reseed_with_int(proof.pow_nonce)
    positions = coin.draw_queries(proof.num_queries, N)
    q = len(positions)
    need(len(proof.base_proofs) == q and len(proof.comp_proofs) == q and len(proof.base_values) == q * nb
         and len(proof.ext_values) == q * ne and len(proof.comp_values) == q * ce, "wrong number of query openings")
    # trace openings
    def check_tree(root, proofs, values, width, what):
        for k, p in enumerate(positions):
            mp, row = proofs[k], values[width * k:width * (k + 1)]
            if width == 1:
                need(mp.variant == 1 and mp.leaf == row[0], f"{what}: opened value differs from the proof's leaf")
                got = root_from_leaf(tree_kind, p, mp.leaf, mp.sibling, mp.path, n_friendly, unhashed=True)
            else:
                need(mp.variant == 0 and mp.leaf == row_digest(tree_kind, row), f"{what}: row hash differs from the proof's leaf")
                got = root_from_leaf(tree_kind, p, mp.leaf, mp.sibling, mp.path, n_friendly)
            need(got == _as_bytes(root), f"{what}: Merkle opening of position {p} does not match the commitment")

    check_tree(proof.base_root, proof.base_proofs, proof.base_values, nb, "base trace")
    if ne:
        check_tree(proof.ext_root, proof.ext_proofs, proof.ext_values, ne, "extension trace")
    check_tree(proof.comp_root, proof.comp_proofs, proof.comp_values, ce, "composition trace")
    # DEEP quotient at the query points (src/lib.rs:102-116: powers of one alpha over the trace arguments, then the composition columns)
    g, w = pow(3, (P - 1) >> log_n, P), pow(3, (P - 1) >> log_N, P)
    zc = pow(z, ce, P)
    deep_at = {}
    for k, p in enumerate(positions):
        x = 3 * pow(w, brev(p, log_N), P) % P
        row = proof.base_values[nb * k:nb * (k + 1)] + proof.ext_values[ne * k:ne * (k + 1)]
        acc, a = 0, 1
        inv = {}
        for (col, off), y in zip(taps, proof.ood_trace):
            if off not in inv:
                inv[off] = pow((x - z * pow(g, off, P)) % P, -1, P)
            acc = (acc + a * (row[col] - y) % P * inv[off]) % P
            a = a * deep_alpha % P
        vinv = pow((x - zc) % P, -1, P)
        for j, y in enumerate(proof.ood_comp):
            acc = (acc + a * (proof.comp_values[ce * k + j] - y) % P * vinv) % P
            a = a * deep_alpha % P
        deep_at[p] = acc
    # FRI
    n_layers = len(proof.fri_layers)
    expect_layers, size = 0, N
    while size // b > proof.fri_max_remainder_coeffs and size > F:
        expect_layers, size = expect_layers + 1, size // F
    need(n_layers == expect_layers and len(proof.remainder_coeffs) == max(1, size // b), "wrong number of FRI layers / remainder coefficients")
    values = dict(deep_at)                                  # entry index -> value, for the current layer
    cur_pos, log_dom, offset = positions, log_N, 3
    for l, layer in enumerate(proof.fri_layers):
        rows = sorted({p >> log_f for p in cur_pos})
        need(len(layer.proofs) == len(rows) and len(layer.flattened_rows) == F * len(rows), f"FRI layer {l}: wrong number of openings")
        nxt = {}
        for k, r in enumerate(rows):
            vals, mp = layer.flattened_rows[F * k:F * (k + 1)], layer.proofs[k]
            need(mp.leaf == row_digest(tree_kind, vals), f"FRI layer {l}: row hash differs from the proof's leaf")
            need(root_from_leaf(tree_kind, r, mp.leaf, mp.sibling, mp.path, n_friendly) == _as_bytes(layer.commitment),
                 f"FRI layer {l}: Merkle opening of row {r} does not match the commitment")
            for j in range(F):
                if F * r + j in values:
                    need(values[F * r + j] == vals[j], f"FRI layer {l}: entry {F * r + j} is not the value folded from the layer above")
            nxt[r] = fri_fold_row(vals, r, log_dom, offset, alphas[l], log_f)
        values, cur_pos, log_dom, offset = nxt, rows, log_dom - log_f, pow(offset, F, P)
    wl = pow(3, (P - 1) >> log_dom, P)
    for r, v in values.items():
        y = offset * pow(wl, brev(r, log_dom), P) % P
        need(remainder_at(proof.remainder_coeffs, y, offset) == v, f"FRI remainder does not match the last fold at entry {r}")
