"""GPU parity: Merkle commitment (all reference tree variants) and Pedersen hash vs the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ss():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import sandstorm_b200

    return sandstorm_b200


KINDS = ["keccak", "keccak_m20", "blake2s_m20", "sha256"]


def kind_ids(ss, oracle, name):
    return {"keccak": (ss.TREE_KECCAK, oracle.TREE_KECCAK), "keccak_m20": (ss.TREE_KECCAK_M20, oracle.TREE_KECCAK_M20),
            "blake2s_m20": (ss.TREE_BLAKE2S_M20, oracle.TREE_BLAKE2S_M20), "sha256": (ss.TREE_SHA256, oracle.TREE_SHA256),
            "friendly": (ss.TREE_FRIENDLY, oracle.TREE_FRIENDLY)}[name]


def check_tree(ss, oracle, cols, name, n_friendly=22, bitrev=False):
    from sandstorm_b200.merkle import MatrixMerkleTree

    gk, ok = kind_ids(ss, oracle, name)
    n = cols.shape[1]
    m = ss.Matrix.from_numpy(cols)
    tree = MatrixMerkleTree.from_matrix(m, gk, n_friendly=n_friendly, row_order=ss.ORDER_BITREV if bitrev else ss.ORDER_NATURAL)
    nodes, leaves, root = oracle.merkle_build(ok, cols, n_friendly=n_friendly, bitrev_rows=bitrev)
    assert tree.root() == root
    assert np.array_equal(tree.leaves(np.arange(n)), leaves)
    assert np.array_equal(tree.nodes(np.arange(1, n)), nodes[1:])
    return tree, nodes, leaves


@pytest.mark.parametrize("name", KINDS)
@pytest.mark.parametrize("n_cols,log_rows", [(1, 1), (1, 6), (2, 1), (2, 3), (3, 7), (7, 10), (9, 11), (8, 9), (10, 5), (17, 4)])
def test_byte_hash_trees_match_oracle(ss, oracle, name, n_cols, log_rows):
    rng = np.random.default_rng(1000 + 31 * n_cols + log_rows)
    cols = oracle.random_felts(rng, n_cols, 1 << log_rows)
    check_tree(ss, oracle, cols, name)


@pytest.mark.parametrize("n_friendly", [0, 1, 2, 3, 22])
@pytest.mark.parametrize("n_cols,log_rows", [(2, 3), (7, 6), (1, 5)])
def test_friendly_trees_match_oracle(ss, oracle, n_friendly, n_cols, log_rows):
    """crypto/src/merkle/mod.rs:505-634 use N_FRIENDLY in {0,1,2,3}; 22 is the production value
    (src/claims.rs:10), for which a 2^6-leaf tree is Pedersen at every level."""
    rng = np.random.default_rng(2000 + n_friendly)
    cols = oracle.random_felts(rng, n_cols, 1 << log_rows)
    check_tree(ss, oracle, cols, "friendly", n_friendly=n_friendly)


def test_bitrev_row_order(ss, oracle):
    rng = np.random.default_rng(3)
    cols = oracle.random_felts(rng, 7, 1 << 8)
    check_tree(ss, oracle, cols, "keccak_m20", bitrev=True)
    check_tree(ss, oracle, cols[:1], "keccak", bitrev=True)


def test_open_and_rows_roundtrip(ss, oracle):
    """Mirror of the reference's prove_rows/verify round trips (crypto/src/merkle/mod.rs:456-634):
    recompute the root from the opened row and its sibling path."""
    rng = np.random.default_rng(4)
    cols = oracle.random_felts(rng, 3, 1 << 6)
    tree, nodes, leaves = check_tree(ss, oracle, cols, "keccak_m20")
    idx = [3, 1, 7, 63, 0]
    proof = tree.prove_rows(idx)
    for k, i in enumerate(idx):
        assert np.array_equal(proof["rows"][k], cols[:, i])
        row_bytes = b"".join(int(sum(int(proof["rows"][k][j][l]) << (64 * l) for l in range(4))).to_bytes(32, "big") for j in range(3))
        h = oracle.hash_bytes(oracle.HASH_KECCAK_M20, row_bytes)
        assert h == bytes(leaves[i])
        pos = i
        for lvl in range(6):
            sib = bytes(proof["paths"][k][lvl])
            h = oracle.hash_bytes(oracle.HASH_KECCAK_M20, (h + sib) if pos % 2 == 0 else (sib + h))
            pos //= 2
        assert h == tree.root()


def test_pedersen_batch_matches_oracle(ss, oracle):
    import ctypes

    import torch

    rng = np.random.default_rng(5)
    n = 300
    a, b = oracle.random_felts(rng, n), oracle.random_felts(rng, n)
    edge = oracle.to_mont([0, 0, oracle.P - 1, 2**248, 1])
    a[:5], b[:5] = edge, edge[::-1]
    da = torch.from_numpy(a.view(np.int64)).cuda()
    db = torch.from_numpy(b.view(np.int64)).cuda()
    out = torch.empty_like(da)
    c = ss.default_context()
    c.check(c.lib.ss_pedersen_hash(c.handle, ctypes.c_void_p(da.data_ptr()), ctypes.c_void_p(db.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), n, None))
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy().view(np.uint64), oracle.pedersen_hash_mont(a, b))
    # reference KAT (builtins/src/pedersen/mod.rs:184-197) through the GPU
    ka = oracle.to_mont([1740729136829561885683894917751815192814966525555656371386868611731128807883])
    kb = oracle.to_mont([919869093895560023824014392670608914007817594969197822578496829435657368346])
    da, db = torch.from_numpy(ka.view(np.int64)).cuda(), torch.from_numpy(kb.view(np.int64)).cuda()
    out = torch.empty_like(da)
    c.check(c.lib.ss_pedersen_hash(c.handle, ctypes.c_void_p(da.data_ptr()), ctypes.c_void_p(db.data_ptr()),
                                   ctypes.c_void_p(out.data_ptr()), 1, None))
    torch.cuda.synchronize()
    assert oracle.from_mont(out.cpu().numpy().view(np.uint64))[0] == 1382171651951541052082654537810074813456022260470662576358627909045455537762


def test_large_tree_root_of_roots(ss, oracle):
    """2^18 rows x 9 columns (starknet base-trace shape): the root must equal the tree built over the
    roots of its four 2^16-row quarter trees (a size-independent consistency property), and a sampled
    quarter must equal the oracle."""
    from sandstorm_b200.merkle import MatrixMerkleTree

    rng = np.random.default_rng(6)
    cols = oracle.random_felts(rng, 9, 1 << 18)
    m = ss.Matrix.from_numpy(cols)
    tree = MatrixMerkleTree.from_matrix(m, ss.TREE_KECCAK_M20)
    quarter_roots = []
    for q in range(4):
        sub = np.ascontiguousarray(cols[:, q << 16:(q + 1) << 16])
        t = MatrixMerkleTree.from_matrix(ss.Matrix.from_numpy(sub), ss.TREE_KECCAK_M20)
        quarter_roots.append(t.root())
        if q == 2:
            assert t.root() == oracle.merkle_build(oracle.TREE_KECCAK_M20, sub)[2]
    H = lambda x, y: oracle.hash_bytes(oracle.HASH_KECCAK_M20, x + y)
    assert tree.root() == H(H(quarter_roots[0], quarter_roots[1]), H(quarter_roots[2], quarter_roots[3]))


@pytest.mark.parametrize("name,n_friendly,n_cols,log_rows", [("keccak_m20", 0, 9, 8), ("friendly", 3, 3, 6), ("friendly", 2, 7, 7), ("friendly", 22, 2, 5)])
def test_sharded_commit_combines_to_the_whole_tree(ss, oracle, name, n_friendly, n_cols, log_rows):
    """The multi-GPU commitment: 4 row-range sub-trees (Pedersen levels counted from the root of the WHOLE tree, so a
    sub-tree keeps n_friendly - 2 of them) + ss_merkle_combine over the 4 sub-roots == the root of the whole tree."""
    import ctypes

    from sandstorm_b200.merkle import MatrixMerkleTree

    gk, ok = kind_ids(ss, oracle, name)
    rng = np.random.default_rng(3000 + log_rows)
    cols = oracle.random_felts(rng, n_cols, 1 << log_rows)
    want = oracle.merkle_build(ok, cols, n_friendly=n_friendly)[2]
    quarter = 1 << (log_rows - 2)
    subs = b""
    for q in range(4):
        sub = np.ascontiguousarray(cols[:, q * quarter:(q + 1) * quarter])
        subs += MatrixMerkleTree.from_matrix(ss.Matrix.from_numpy(sub), gk, n_friendly=max(0, n_friendly - 2)).root()
    c = ss.default_context()
    root = (ctypes.c_uint8 * 32)()
    buf = (ctypes.c_uint8 * 128).from_buffer_copy(subs)
    c.check(c.lib.ss_merkle_combine(c.handle, gk, buf, 2, root))
    assert bytes(root) == want


@pytest.mark.parametrize("name,n_friendly", [("keccak_m20", 0), ("friendly", 3), ("friendly", 22)])
def test_leaf_ranges_trees_from_leaves_and_fri_row_order(ss, oracle, name, n_friendly):
    """The pieces a multi-GPU commitment is made of: digests of ranges of tree leaves (ss_hash_rows) in the three orders,
    a tree over digests computed elsewhere (ss_merkle_build_from_leaves), the 32-byte bit-reversal permutation, and the FRI
    layer order (rows and columns bit-reversed) — against the oracle's tree of the explicitly reordered matrix."""
    import ctypes

    import torch

    from sandstorm_b200 import _lib

    gk, ok = kind_ids(ss, oracle, name)
    rng = np.random.default_rng(77)
    log_rows, n_cols = 11, 8
    n = 1 << log_rows
    cols = oracle.random_felts(rng, n_cols, n)
    m = ss.Matrix.from_numpy(cols)
    c = m.ctx
    perm = [int(f"{j:03b}"[::-1], 2) for j in range(n_cols)]
    for order, ref_cols, ref_bitrev in ((_lib.ORDER_NATURAL, cols, False), (_lib.ORDER_BITREV, cols, True), (_lib.ORDER_BITREV_RC, cols[perm], True)):
        nodes, leaves, root = oracle.merkle_build(ok, np.ascontiguousarray(ref_cols), n_friendly=n_friendly, bitrev_rows=ref_bitrev)
        got = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
        for lo, cnt in ((0, 5), (5, 1019), (1024, 1024)):
            c.check(c.lib.ss_hash_rows(c.handle, gk, ctypes.c_void_p(m.data.data_ptr()), m.col_stride, n_cols, log_rows, order, lo, cnt,
                                       ctypes.c_void_p(got[lo].data_ptr()), None))
        torch.cuda.synchronize()
        assert np.array_equal(got.cpu().numpy().view(np.uint8).reshape(n, 32), leaves)
        handle = ctypes.c_void_p()
        c.check(c.lib.ss_merkle_build_from_leaves(c.handle, gk, n_friendly, ctypes.c_void_p(got.data_ptr()), log_rows, ctypes.byref(handle), None))
        out = (ctypes.c_uint8 * 32)()
        c.check(c.lib.ss_merkle_root(c.handle, handle, out))
        c.lib.ss_tree_free(handle)
        assert bytes(out) == root
        # the whole-matrix entry point in the same order
        from sandstorm_b200.merkle import MatrixMerkleTree

        assert MatrixMerkleTree.from_matrix(m, gk, n_friendly=n_friendly, row_order=order).root() == root
    # natural-order digests -> tree order by the in-place permutation
    nat = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    c.check(c.lib.ss_hash_rows(c.handle, gk, ctypes.c_void_p(m.data.data_ptr()), m.col_stride, n_cols, log_rows, _lib.ORDER_NATURAL, 0, n,
                               ctypes.c_void_p(nat.data_ptr()), None))
    c.check(c.lib.ss_bitrev_permute32(c.handle, ctypes.c_void_p(nat.data_ptr()), log_rows, None))
    want = torch.zeros((n, 4), dtype=torch.int64, device="cuda")
    c.check(c.lib.ss_hash_rows(c.handle, gk, ctypes.c_void_p(m.data.data_ptr()), m.col_stride, n_cols, log_rows, _lib.ORDER_BITREV, 0, n,
                               ctypes.c_void_p(want.data_ptr()), None))
    torch.cuda.synchronize()
    assert torch.equal(nat, want)
