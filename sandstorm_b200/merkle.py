"""Mirror of the reference's ``MatrixMerkleTree`` / ``MerkleTree`` trait surface for the GPU path
(crypto/src/merkle/mod.rs:64-166 FriendlyMerkleTree, :254-347 LeafVariantMerkleTree):

    tree = MatrixMerkleTree.from_matrix(matrix, kind)      # from_matrix
    tree.root()                                            # MerkleTree::root (Digest::as_bytes)
    proof = tree.prove_rows(indices)                       # MerkleTree::prove + row values
    MatrixMerkleTree.verify_rows(kind, root, indices, rows, proof)   # verify (host-side check)

The tree lives on the device (``ss_tree``) until the object is dropped.
"""
from __future__ import annotations

import ctypes
import hashlib

import numpy as np
import torch

from . import _lib
from .matrix import Matrix, _stream_ptr

NUM_FRIENDLY_COMMITMENT_LAYERS = 22     # src/claims.rs:10


class MerkleError(Exception):
    """ministark::merkle::Error::InvalidProof"""


class MatrixMerkleTree:
    def __init__(self, handle, matrix: Matrix, kind: int, n_friendly: int, row_order: int):
        self._handle = handle
        self.matrix = matrix
        self.ctx = matrix.ctx
        self.kind = kind
        self.n_friendly = n_friendly
        self.row_order = row_order
        self.log_rows = matrix.log_rows

    @classmethod
    def from_matrix(cls, matrix: Matrix, kind: int, n_friendly: int = NUM_FRIENDLY_COMMITMENT_LAYERS,
                    row_order: int = _lib.ORDER_NATURAL) -> "MatrixMerkleTree":
        c = matrix.ctx
        handle = ctypes.c_void_p()
        c.check(c.lib.ss_merkle_build(c.handle, kind, n_friendly, ctypes.c_void_p(matrix.data.data_ptr()), matrix.col_stride,
                                      matrix.num_cols, matrix.log_rows, row_order, ctypes.byref(handle), _stream_ptr()))
        return cls(handle, matrix, kind, n_friendly, row_order)

    def __del__(self):
        try:
            if self._handle:
                self.ctx.lib.ss_tree_free(self._handle)
                self._handle = None
        except Exception:
            pass

    # ---- MerkleTree ---------------------------------------------------------------------------
    def root(self) -> bytes:
        out = (ctypes.c_uint8 * 32)()
        self.ctx.check(self.ctx.lib.ss_merkle_root(self.ctx.handle, self._handle, out))
        return bytes(out)

    def _gather(self, fn, indices) -> np.ndarray:
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        out = np.zeros((len(idx), 32), dtype=np.uint8)
        self.ctx.check(fn(self.ctx.handle, self._handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        return out

    def nodes(self, indices) -> np.ndarray:
        return self._gather(self.ctx.lib.ss_merkle_nodes, indices)

    def leaves(self, indices) -> np.ndarray:
        return self._gather(self.ctx.lib.ss_merkle_leaves, indices)

    def prove(self, indices) -> np.ndarray:
        """Sibling paths, leaf level first: uint8[len(indices), log_rows, 32] (storage form)."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        out = np.zeros((len(idx), self.log_rows, 32), dtype=np.uint8)
        self.ctx.check(self.ctx.lib.ss_merkle_open(self.ctx.handle, self._handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                                   len(idx), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
        return out

    def rows(self, indices) -> np.ndarray:
        """Matrix::read_row for each committed leaf index: uint64[len, n_cols, 4]."""
        idx = np.ascontiguousarray(indices, dtype=np.uint64)
        if self.row_order == _lib.ORDER_BITREV:
            idx = np.array([int(f"{int(i):0{self.log_rows}b}"[::-1], 2) for i in idx], dtype=np.uint64)
        m = self.matrix
        out = np.zeros((len(idx), m.num_cols, 4), dtype=np.uint64)
        self.ctx.check(self.ctx.lib.ss_rows_gather(self.ctx.handle, ctypes.c_void_p(m.data.data_ptr()), m.col_stride, m.num_cols,
                                                   idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                                   out.ctypes.data_as(ctypes.c_void_p)))
        return out

    def prove_rows(self, indices):
        return {"rows": self.rows(indices), "paths": self.prove(indices)}

    # ---- MerkleTree::verify / MatrixMerkleTree::verify_rows (host side, as in the reference: crypto/src/merkle/mod.rs:125-165, 306-346) ----
    @staticmethod
    def verify_rows(kind: int, root: bytes, indices, rows: np.ndarray, paths: np.ndarray, n_friendly: int = NUM_FRIENDLY_COMMITMENT_LAYERS) -> None:
        """Recomputes the root from the opened rows (uint64[q, n_cols, 4] Montgomery limbs) and their sibling paths
        (uint8[q, depth, 32], leaf level first, storage form); raises MerkleError on a mismatch."""
        from .verify import merkle_root_from_opening

        for q, idx in enumerate(indices):
            if merkle_root_from_opening(kind, int(idx), rows[q], paths[q], n_friendly) != root:
                raise MerkleError(f"invalid Merkle opening of row {int(idx)}")
