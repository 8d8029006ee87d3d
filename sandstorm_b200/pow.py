"""Proof-of-work grinding on the GPU: `PublicCoin::grind_proof_of_work` of the reference's coins
(crypto/src/public_coin/solidity.rs:120-141 — Keccak-256; cairo.rs:133-154 — Blake2s-256).  Returns the smallest valid
nonce (the reference's sequential answer; its `-F parallel` build returns any valid nonce)."""
from __future__ import annotations

import ctypes

from .context import Context, default_context

POW_KECCAK, POW_BLAKE2S = 0, 1


def grind_proof_of_work(digest: bytes, bits: int, hash_kind: int = POW_KECCAK, ctx: Context | None = None) -> int:
    if len(digest) != 32:
        raise ValueError("digest must be 32 bytes")
    c = ctx or default_context()
    out = ctypes.c_uint64()
    buf = (ctypes.c_uint8 * 32).from_buffer_copy(digest)
    c.check(c.lib.ss_pow_grind(c.handle, hash_kind, buf, bits, ctypes.byref(out)))
    return out.value
