"""sandstorm_b200 — B200 (sm_100a) backend for the hot path of `sandstorm prove`.

Host-side mirror of the reference's trait surface for that path:
``Matrix`` (ministark::Matrix: interpolate / evaluate), ``MatrixMerkleTree.from_matrix`` with the
reference's tree variants (crypto/src/merkle/mod.rs), FRI folding and constraint evaluation, all
running through the C ABI in include/sandstorm_b200.h.  PyTorch is used only for device memory,
streams and torch.distributed plumbing.
"""
from ._lib import (FIELD_FP252, FIELD_GOLDILOCKS, ORDER_BITREV, ORDER_NATURAL, TREE_BLAKE2S_M20, TREE_FRIENDLY,
                   TREE_KECCAK, TREE_KECCAK_M20, TREE_SHA256, SandstormError)
from .context import Context, default_context
from .matrix import Matrix

__all__ = [
    "Context", "default_context", "Matrix", "SandstormError",
    "FIELD_FP252", "FIELD_GOLDILOCKS", "ORDER_NATURAL", "ORDER_BITREV",
    "TREE_KECCAK", "TREE_KECCAK_M20", "TREE_FRIENDLY", "TREE_BLAKE2S_M20", "TREE_SHA256",
]
