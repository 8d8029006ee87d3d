"""The two proofs the reference ships (bootloader-proof.bin: recursive layout, trace 2^18, 40 queries;
example/array-sum.proof.saved: starknet layout, trace 2^21, 16 queries — both masked-Keccak LeafVariant trees), copied
to tests/golden/reference_proofs/.  They are the only reference-MADE artefacts of the hot path in the tree, and they pin:

  * the `Proof` wire format (sandstorm_b200/proof.py): parsed to the last byte, re-serialized byte for byte;
  * row hashing, node hashing / masking, tree layout and path order (sandstorm_b200/verify.py, the same conventions as
    csrc/merkle.cu and oracle/merkle.c): every opening of every tree recomputes the committed root;
  * bit-reversed commitment order, the FRI layer orientation, the fold formula (no 1/F) and the remainder convention: for
    every layer ONE alpha (solved from two queries as the common root of their fold equations) satisfies all queries.
What needs the public coin (positions, alphas, OOD point) cannot be recomputed: the runs' public inputs are not in the tree."""
import os

import pytest

from sandstorm_b200 import _lib
from sandstorm_b200.proof import HASHED, UNHASHED, Proof
from sandstorm_b200.verify import P, brev, fri_fold_row, remainder_at, root_from_leaf, row_digest

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_proofs")
CASES = [("recursive_keccak_2p18.proof.bin", 7, 3), ("starknet_keccak_2p21.proof.bin", 9, 1)]
KIND = _lib.TREE_KECCAK_M20


@pytest.fixture(autouse=True)
def fast_keccak(monkeypatch, oracle):
    """the pure-Python Keccak of sandstorm_b200/hostcrypto.py is checked against the C one in tests/test_public_coin.py;
    the thousands of hashes below use the C one."""
    from sandstorm_b200 import hostcrypto as hc

    monkeypatch.setattr(hc, "keccak256", lambda d: oracle.hash_bytes(oracle.HASH_KECCAK, bytes(d)))


@pytest.fixture(scope="module", params=CASES, ids=[c[0] for c in CASES])
def case(request):
    name, n_base, n_ext = request.param
    data = open(os.path.join(HERE, name), "rb").read()
    return data, Proof.deserialize(data), n_base, n_ext


def test_wire_format_round_trip(case):
    data, proof, n_base, n_ext = case
    assert proof.serialize() == data
    q = len(proof.base_proofs)
    assert (proof.lde_blowup_factor, proof.grinding_factor, proof.fri_folding_factor) == (2, 16, 8) and q <= proof.num_queries
    assert len(proof.base_values) == q * n_base and len(proof.ext_values) == q * n_ext and len(proof.comp_values) == q * 2
    log_N = (proof.trace_len * 2).bit_length() - 1
    assert all(len(p.path) == log_N - 1 for p in proof.base_proofs + proof.ext_proofs + proof.comp_proofs)
    assert [len(l.proofs[0].path) for l in proof.fri_layers] == [log_N - 3 * (k + 1) - 1 for k in range(len(proof.fri_layers))]
    assert 0 < len(proof.remainder_coeffs) <= proof.fri_max_remainder_coeffs and len(proof.ood_comp) == 2
    assert len(proof.ood_trace) == {7: 133, 9: 269}[n_base]                 # the mask sizes of the two layouts


def positions_of_layer(oracle, layer):
    out = []
    for p in layer.proofs:
        idx = oracle.merkle_find_index(oracle.HASH_KECCAK_M20, p.leaf, [p.sibling] + p.path, layer.commitment)
        assert idx >= 0
        out.append(idx)
    return out


def test_every_opening_recomputes_its_root(case, oracle):
    data, proof, n_base, n_ext = case
    q = len(proof.base_proofs)
    fold = proof.fri_folding_factor
    # FRI layer 0: leaf digests are the hashes of the flattened rows; positions recovered by search in the (small) layer trees
    rows0 = positions_of_layer(oracle, proof.fri_layers[0])
    assert rows0 == sorted(set(rows0))
    for k, p in enumerate(proof.fri_layers[0].proofs):
        assert p.variant == HASHED and p.leaf == row_digest(KIND, proof.fri_layers[0].flattened_rows[fold * k:fold * (k + 1)])
    # trace trees: query k sits in FRI row position >> 3 -> eight candidate positions, exactly one verifies; the SAME position
    # must then verify in the extension and composition trees
    positions = []
    for k in range(q):
        bp = proof.base_proofs[k]
        assert bp.leaf == row_digest(KIND, proof.base_values[n_base * k:n_base * (k + 1)])
        # (positions ascend, so query k lies in row k - (number of duplicates so far) or the one before / after)
        near = rows0[max(0, k - (q - len(rows0)) - 1):k + 1]
        hits = [8 * r + j for r in near for j in range(8) if root_from_leaf(KIND, 8 * r + j, bp.leaf, bp.sibling, bp.path) == proof.base_root]
        assert len(hits) == 1
        positions.append(hits[0])
    assert positions == sorted(set(positions)) and sorted({p >> 3 for p in positions}) == rows0
    for k, pos in enumerate(positions):
        ep, cp = proof.ext_proofs[k], proof.comp_proofs[k]
        if n_ext == 1:
            assert ep.variant == UNHASHED and ep.leaf == proof.ext_values[k]
            assert root_from_leaf(KIND, pos, ep.leaf, ep.sibling, ep.path, unhashed=True) == proof.ext_root
        else:
            assert ep.leaf == row_digest(KIND, proof.ext_values[n_ext * k:n_ext * (k + 1)])
            assert root_from_leaf(KIND, pos, ep.leaf, ep.sibling, ep.path) == proof.ext_root
        assert cp.leaf == row_digest(KIND, proof.comp_values[2 * k:2 * k + 2])
        assert root_from_leaf(KIND, pos, cp.leaf, cp.sibling, cp.path) == proof.comp_root
    # deeper FRI layers: positions fold by >> 3
    prev = rows0
    for layer in proof.fri_layers[1:]:
        cur = positions_of_layer(oracle, layer)
        assert cur == sorted({r >> 3 for r in prev})
        prev = cur


def poly_gcd(a, b):
    def trim(p):
        while p and p[-1] == 0:
            p.pop()
        return p

    a, b = trim(a[:]), trim(b[:])
    while b:
        inv = pow(b[-1], -1, P)
        r = a[:]
        while len(r) >= len(b) and r:
            c = r[-1] * inv % P
            for i in range(len(b)):
                r[len(r) - len(b) + i] = (r[len(r) - len(b) + i] - c * b[i]) % P
            trim(r)
        a, b = b, r
    return a


def test_fri_folds_are_consistent_for_one_alpha_per_layer(case, oracle):
    data, proof, n_base, n_ext = case
    log_N = (proof.trace_len * 2).bit_length() - 1
    pos = [positions_of_layer(oracle, layer) for layer in proof.fri_layers]
    n_layers = len(proof.fri_layers)
    for l, layer in enumerate(proof.fri_layers):
        log_dom, offset = log_N - 3 * l, pow(3, 8 ** l, P)
        last = l == n_layers - 1
        # fold value as a polynomial in alpha: coefficient m = fold with alpha^m isolated (fold is linear in the powers of alpha)
        polys = []
        for k, r in enumerate(pos[l]):
            vals = layer.flattened_rows[8 * k:8 * k + 8]
            basis = [(fri_fold_row(vals, r, log_dom, offset, 0) if m == 0 else 0) for m in range(8)]
            w = pow(3, (P - 1) >> log_dom, P)
            x_inv = pow(offset * pow(w, brev(r, log_dom - 3), P) % P, -1, P)
            wF_inv = pow(3, -((P - 1) >> 3), P)
            f = [vals[brev(j, 3)] for j in range(8)]
            coeffs = [pow(x_inv, m, P) * sum(f[j] * pow(wF_inv, m * j, P) for j in range(8)) % P for m in range(8)]
            assert coeffs[0] == basis[0]
            if last:
                y = pow(offset * pow(w, brev(r, log_dom - 3), P) % P, 8, P)
                target = remainder_at(proof.remainder_coeffs, y, pow(offset, 8, P))
            else:
                nxt = proof.fri_layers[l + 1]
                j = pos[l + 1].index(r >> 3)
                target = nxt.flattened_rows[8 * j + (r & 7)]
            coeffs[0] = (coeffs[0] - target) % P
            polys.append(coeffs)
        g = poly_gcd(polys[0], polys[1])
        assert len(g) == 2, f"layer {l}: the fold equations of two queries have no single common root"
        alpha = -g[0] * pow(g[1], -1, P) % P
        for k, r in enumerate(pos[l]):
            vals = layer.flattened_rows[8 * k:8 * k + 8]
            got = fri_fold_row(vals, r, log_dom, offset, alpha)
            want = polys[k][0]                                     # (coefficient 0 holds fold_0 - target)
            assert sum(c * pow(alpha, m, P) for m, c in enumerate(polys[k])) % P == 0
            assert got == (sum(c * pow(alpha, m, P) for m, c in enumerate(polys[k])) + got) % P and want is not None
