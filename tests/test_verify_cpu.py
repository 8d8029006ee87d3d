"""Host-side Merkle verification (sandstorm_b200/verify.py: MerkleTree::verify / verify_rows of every tree variant)
against trees built by the C oracle: the opening of any leaf recomputes the oracle's root; a corrupted row or path does not."""
import numpy as np
import pytest

from sandstorm_b200 import _lib
from sandstorm_b200.verify import merkle_root_from_opening


def opening(nodes, leaves, log_rows, idx):
    """sibling path, leaf level first, from the oracle's node / leaf arrays (the layout ss_merkle_open reads)."""
    n = 1 << log_rows
    path = [leaves[idx ^ 1]]
    node = (n + idx) >> 1
    while node > 1:
        path.append(nodes[node ^ 1])
        node >>= 1
    return np.array(path, dtype=np.uint8)


@pytest.mark.parametrize("kind,n_cols,n_friendly", [("KECCAK_M20", 3, 0), ("KECCAK", 2, 0), ("KECCAK_M20", 1, 0), ("BLAKE2S_M20", 4, 0), ("SHA256", 2, 0),
                                                    ("FRIENDLY", 3, 22), ("FRIENDLY", 3, 2), ("FRIENDLY", 2, 0), ("FRIENDLY", 1, 22)])
def test_openings_recompute_the_root(oracle, kind, n_cols, n_friendly):
    log_rows = 4
    rng = np.random.default_rng(n_cols * 7 + n_friendly)
    cols = oracle.random_felts(rng, n_cols, 1 << log_rows)
    ok, gk = getattr(oracle, "TREE_" + kind), getattr(_lib, "TREE_" + kind)
    nodes, leaves, root = oracle.merkle_build(ok, cols, n_friendly)
    nodes, leaves = np.frombuffer(nodes, dtype=np.uint8).reshape(-1, 32), np.frombuffer(leaves, dtype=np.uint8).reshape(-1, 32)
    for idx in (0, 5, 10, 15):
        path = opening(nodes, leaves, log_rows, idx)
        row = cols[:, idx]
        assert merkle_root_from_opening(gk, idx, row, path, n_friendly) == root
        bad = row.copy()
        bad[0, 0] ^= 1
        assert merkle_root_from_opening(gk, idx, bad, path, n_friendly) != root
        bad_path = path.copy()
        bad_path[-1, 3] ^= 1
        assert merkle_root_from_opening(gk, idx, row, bad_path, n_friendly) != root
