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
