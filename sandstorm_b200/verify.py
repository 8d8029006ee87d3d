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
