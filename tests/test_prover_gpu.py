"""GPU: the whole device hot path (sandstorm_b200/prover.py) on the recursive layout — commitments equal
the oracle's, OOD values equal Horner evaluations, and the FRI remainder is a low-degree polynomial
(which exercises constraint evaluation -> composition split -> DEEP quotient -> folds end to end)."""
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


@pytest.mark.parametrize("layout,log_n,tree", [("plain", 7, "keccak_m20"), ("recursive", 12, "keccak_m20"), ("recursive", 11, "friendly"),
                                               ("starknet", 16, "keccak_m20")])
def test_hot_path_is_consistent(ss, oracle, layout, log_n, tree):
    import torch

    from sandstorm_b200.prover import HotPathProver, ProofOptions

    P = oracle.P
    # recursive claims commit with FriendlyMerkleTree<22, Blake2s masked / Pedersen> (src/claims.rs:10,29-32)
    gk, ok = (ss.TREE_FRIENDLY, oracle.TREE_FRIENDLY) if tree == "friendly" else (ss.TREE_KECCAK_M20, oracle.TREE_KECCAK_M20)
    hp = HotPathProver(layout, log_n, ProofOptions(num_queries=12, tree_kind=gk))
    L = hp.layout
    rng = np.random.default_rng(log_n)
    base = oracle.random_felts(rng, L.num_base_columns, 1 << log_n)
    ext = oracle.random_felts(rng, L.num_extension_columns, 1 << log_n)
    res = hp.prove(ss.Matrix.from_numpy(base), ss.Matrix.from_numpy(ext), self_check=True, keep_openings=True)
    torch.cuda.synchronize()
    # the DEEP quotient, evaluated on the sub-coset 3<w_n> and extended, equals its evaluation on every LDE row
    assert res.deep_matches_full_evaluation is True
    # commitments
    # (rows are committed in bit-reversed order of the LDE domain: the reference's convention, tests/test_reference_proof.py)
    assert res.roots["base"] == oracle.merkle_build(ok, oracle.lde(base, 1), bitrev_rows=True)[2]
    assert res.roots["ext"] == oracle.merkle_build(ok, oracle.lde(ext, 1), bitrev_rows=True)[2]
    # out-of-domain values: every tap of the mask against Horner evaluation of the oracle's interpolation at z * g^offset
    coeffs = [oracle.from_mont(c) for c in oracle.ntt(np.concatenate([base, ext]), inverse=True)]
    taps = L.taps()
    g, z = pow(3, (P - 1) >> log_n, P), res.ood_point
    assert len(res.ood_trace) == len(taps) and len(res.ood_composition) == 2
    for k in sorted({0, 1, len(taps) // 2, len(taps) - 1}):
        col, off = taps[k]
        x, acc = z * pow(g, off, P) % P, 0
        for v in reversed(coeffs[col]):
            acc = (acc * x + v) % P
        assert res.ood_trace[k] == acc, (k, col, off)
    # the composition columns: commitment of their LDE, and the opened rows at the query positions are rows of that LDE
    comp_rows = res.trace_queries["composition"]["rows"]
    assert comp_rows.shape == (len(res.query_positions), 2, 4)
    # FRI remainder: the coefficients of f(offset * X) for the last layer's codeword, whose upper half vanished
    log_m, offset = hp.final_domain
    rem = oracle.from_mont(res.remainder)
    assert len(rem) == (1 << log_m) >> 1 and any(rem) and hp.remainder_high_zero is True
    assert len(res.fri_roots) >= 1 and res.opened_bytes > 0
    # opened base rows: leaf p commits the LDE row brev(p)
    lde_base = oracle.lde(base, 1)
    bits = log_n + 1
    for q, pos in enumerate(res.query_positions[:4]):
        assert np.array_equal(res.trace_queries["base"]["rows"][q], lde_base[:, int(f"{pos:0{bits}b}"[::-1], 2)])
    # every opening verifies on the host against the committed root (MatrixMerkleTree::verify_rows)
    from sandstorm_b200.merkle import MatrixMerkleTree

    for name in ("base", "ext", "composition"):
        tq = res.trace_queries[name]
        MatrixMerkleTree.verify_rows(gk, res.roots[name], res.query_positions[:3], tq["rows"][:3], tq["paths"][:3])


@pytest.mark.parametrize("layout,log_n,tree", [("plain", 7, "keccak_m20"), ("recursive", 12, "friendly"), ("starknet", 16, "keccak_m20")])
def test_gpu_pipeline_equals_cpu_pipeline(ss, oracle, layout, log_n, tree):
    """The whole hot path on the GPU against the whole hot path on the CPU (oracle/prover.py: Horner OOD, DEEP quotient on
    every row from its definition, FRI folds from the definition) with the same coin: every commitment — including the
    composition root, i.e. constraint evaluation + split — every out-of-domain value, FRI root and the remainder bit for bit."""
    import torch

    from oracle.prover import CpuHotPath
    from sandstorm_b200.prover import HotPathProver, ProofOptions, SeededCoin

    gk, ok = (ss.TREE_FRIENDLY, oracle.TREE_FRIENDLY) if tree == "friendly" else (ss.TREE_KECCAK_M20, oracle.TREE_KECCAK_M20)
    hp = HotPathProver(layout, log_n, ProofOptions(num_queries=8, tree_kind=gk), coin=SeededCoin(77))
    L = hp.layout
    rng = np.random.default_rng(1000 + log_n)
    base = oracle.random_felts(rng, L.num_base_columns, 1 << log_n)
    ext = oracle.random_felts(rng, L.num_extension_columns, 1 << log_n)
    got = hp.prove(ss.Matrix.from_numpy(base), ss.Matrix.from_numpy(ext), queries=False)
    torch.cuda.synchronize()
    want = CpuHotPath(layout, log_n, tree_kind=ok).prove(base, ext, SeededCoin(77))
    assert got.roots == want["roots"]
    assert got.ood_trace == want["ood_trace"] and got.ood_composition == want["ood_composition"]
    assert got.fri_roots == want["fri_roots"] and len(got.fri_roots) >= 1
    assert np.array_equal(got.remainder, want["remainder"]) and want["remainder_high_zero"] and hp.remainder_high_zero


def test_ood_values_match_horner(ss, oracle):
    """poly_eval in both coefficient formats against Horner on the oracle's interpolation."""
    from sandstorm_b200.matrix import poly_eval

    P = oracle.P
    rng = np.random.default_rng(5)
    cols = oracle.random_felts(rng, 2, 1 << 9)
    lde, coeffs = ss.Matrix.from_numpy(cols).lde(1, keep_coeffs=True)
    plain = oracle.ntt(cols, inverse=True)
    z = int.from_bytes(rng.bytes(31), "big")
    want = []
    for c in range(2):
        acc = 0
        for v in reversed(oracle.from_mont(plain[c])):
            acc = (acc * z + v) % P
        want.append(acc)
    assert oracle.from_mont(poly_eval(coeffs, [0, 1], oracle.to_mont([z, z]))) == want
    assert oracle.from_mont(poly_eval(ss.Matrix.from_numpy(plain), [0, 1], oracle.to_mont([z, z]), natural_order=True)) == want


@pytest.mark.parametrize("log_n", [3, 9, 13])
def test_barycentric_ood_matches_horner(ss, oracle, log_n):
    """ss_ood_eval (dot products of the trace with the barycentric weights) against Horner evaluation of the
    oracle's interpolation, for wrapped and unwrapped row offsets, whole domain and row ranges that add up."""
    from sandstorm_b200.matrix import ood_eval

    P = oracle.P
    n = 1 << log_n
    rng = np.random.default_rng(60 + log_n)
    cols = oracle.random_felts(rng, 3, n)
    coeffs = [oracle.from_mont(c) for c in oracle.ntt(cols, inverse=True)]
    z = int.from_bytes(rng.bytes(31), "big")
    g = pow(3, (P - 1) >> log_n, P)
    taps = [(0, 0), (0, 1), (0, n - 1), (1, 2 % n), (1, n // 2), (2, 5 % n), (2, 0), (2, 1), (2, 3 % n), (0, n + 1)]

    def horner(c, x):
        acc = 0
        for v in reversed(coeffs[c]):
            acc = (acc * x + v) % P
        return acc

    want = [horner(c, z * pow(g, off, P) % P) for c, off in taps]
    m = ss.Matrix.from_numpy(cols)
    zm = oracle.to_mont([z])[0]
    got = oracle.from_mont(ood_eval(m, [c for c, _ in taps], [off for _, off in taps], zm))
    assert got == want
    if n >= 8:
        cuts = [0, n // 8, n // 2 + 1, n]
        parts = [oracle.from_mont(ood_eval(m, [c for c, _ in taps], [off for _, off in taps], zm, rows=(a, b - a))) for a, b in zip(cuts, cuts[1:])]
        assert [sum(v) % P for v in zip(*parts)] == want


def test_ood_values_do_not_depend_on_how_they_are_computed(ss, oracle):
    """out-of-domain mask values from per-tap barycentric sums (ood_transform_min_taps=0) == from one coset transform per
    column (ss_coset_eval; min_taps=1 sends every column that way) == Horner on the oracle's interpolation."""
    import torch

    from sandstorm_b200.prover import HotPathProver, ProofOptions

    P, log_n = oracle.P, 11
    rng = np.random.default_rng(5)
    got = {}
    for min_taps in (0, 1, 20):
        hp = HotPathProver("recursive", log_n, ProofOptions(num_queries=4, ood_transform_min_taps=min_taps))
        L = hp.layout
        if not got:
            base = oracle.random_felts(rng, L.num_base_columns, 1 << log_n)
            ext = oracle.random_felts(rng, L.num_extension_columns, 1 << log_n)
        res = hp.prove(ss.Matrix.from_numpy(base), ss.Matrix.from_numpy(ext), queries=False)
        torch.cuda.synchronize()
        got[min_taps] = (res.ood_point, list(res.ood_trace), res.fri_roots)
    assert got[0] == got[1] == got[20]
    coeffs = [oracle.from_mont(c) for c in oracle.ntt(np.concatenate([base, ext]), inverse=True)]
    g, z = pow(3, (P - 1) >> log_n, P), got[1][0]
    for k, (col, off) in enumerate(L.taps()):
        if k % 7 == 0:
            assert got[1][1][k] == sum(v * pow(z * pow(g, off, P) % P, e, P) for e, v in enumerate(coeffs[col])) % P


@pytest.mark.parametrize("layout,log_n", [("recursive", 11), ("starknet", 16)])
def test_deep_quotient_does_not_depend_on_how_its_pole_sums_are_computed(ss, oracle, layout, log_n):
    """DEEP quotient with every pole a shifted read of u (deep_filter_min_taps=0) == with V and the tap-heavy columns' W_c taken from
    two transforms each (default) == with EVERY column's W_c from transforms (min_taps=1): same FRI roots and remainder; and the
    extended quotient equals the per-row form on every LDE row (self_check)."""
    import torch

    from sandstorm_b200.prover import HotPathProver, ProofOptions

    rng = np.random.default_rng(9)
    got = {}
    for min_taps in (0, 40, 1):
        hp = HotPathProver(layout, log_n, ProofOptions(num_queries=4, deep_filter_min_taps=min_taps))
        L = hp.layout
        if not got:
            base = oracle.random_felts(rng, L.num_base_columns, 1 << log_n)
            ext = oracle.random_felts(rng, L.num_extension_columns, 1 << log_n)
        assert (hp.value_col is None) == (min_taps == 0)
        res = hp.prove(ss.Matrix.from_numpy(base), ss.Matrix.from_numpy(ext), queries=False, self_check=True)
        torch.cuda.synchronize()
        assert res.deep_matches_full_evaluation is True
        got[min_taps] = (res.fri_roots, res.remainder.tobytes())
    assert got[0] == got[40] == got[1]
