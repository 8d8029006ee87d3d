"""Pins the CPU oracle to every known-answer test the reference holds for the hot path
(SURVEY.md §8c).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest

import ecref
from ecref import P

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load_coeffs():
    with open(os.path.join(GOLDEN, "periodic_coeffs.json")) as f:
        return {k: [int(v, 16) for v in vals] for k, vals in json.load(f).items()}


def doubling_table(pt, count):
    out = []
    for _ in range(count):
        out.append(pt)
        pt = ecref.ec_add(pt, pt)
    return out


# ------------------------------------------------------------------ field core
def test_fp_mul_matches_bigint(oracle):
    rng = np.random.default_rng(1)
    a, b = oracle.random_felts(rng, 200), oracle.random_felts(rng, 200)
    got = oracle.from_mont(oracle.fp_mul(a, b))
    want = [x * y % P for x, y in zip(oracle.from_mont(a), oracle.from_mont(b))]
    assert got == want
    edge = oracle.to_mont([0, 1, P - 1, P - 2, 2**251, 3])
    for i in range(len(edge)):
        for j in range(len(edge)):
            x, y = oracle.from_mont(edge[i:i + 1])[0], oracle.from_mont(edge[j:j + 1])[0]
            assert oracle.from_mont(oracle.fp_mul(edge[i:i + 1], edge[j:j + 1]))[0] == x * y % P


def test_montgomery_r_is_2_256(oracle):
    # MONTGOMERY_R quoted at crypto/src/utils.rs:20-21
    one = oracle.to_mont([1])[0]
    val = sum(int(one[i]) << (64 * i) for i in range(4))
    assert val == 3618502788666127798953978732740734578953660990361066340291730267701097005025


# -------------------------------------------------------------------- NTT KATs
def test_ntt_small_matches_naive_dft(oracle):
    rng = np.random.default_rng(2)
    for log_n in range(1, 7):
        cols = oracle.random_felts(rng, 2, 1 << log_n)
        got = oracle.ntt(cols)
        for c in range(2):
            assert oracle.from_mont(got[c]) == ecref.naive_dft(oracle.from_mont(cols[c]))
        back = oracle.ntt(got, inverse=True)
        assert np.array_equal(back, cols)


def test_pedersen_periodic_columns_kat(oracle):
    """builtins/src/pedersen/periodic.rs:1184-1209: fft(HASH_POINTS_{X,Y}_COEFFS) == EC doubling table."""
    coeffs = load_coeffs()
    pts = []
    for base_lo, base_hi in ((1, 2), (3, 4)):
        part = doubling_table(ecref.PEDERSEN_P[base_lo], 248) + doubling_table(ecref.PEDERSEN_P[base_hi], 4)
        part += [part[-1]] * 4
        pts += part
    assert len(pts) == 512
    for axis, name in enumerate(("HASH_POINTS_X_COEFFS", "HASH_POINTS_Y_COEFFS")):
        evals = oracle.from_mont(oracle.ntt(oracle.to_mont(coeffs[name])[None])[0])
        assert evals == [pt[axis] for pt in pts]


def test_ecdsa_periodic_columns_kat(oracle):
    """builtins/src/ecdsa/periodic.rs:599-625."""
    coeffs = load_coeffs()
    pts = doubling_table(ecref.EC_GENERATOR, 251)
    pts += [pts[-1]] * 5
    for axis, name in enumerate(("GENERATOR_POINTS_X_COEFFS", "GENERATOR_POINTS_Y_COEFFS")):
        evals = oracle.from_mont(oracle.ntt(oracle.to_mont(coeffs[name])[None])[0])
        assert evals == [pt[axis] for pt in pts]


def test_poseidon_full_round_keys_kat(oracle):
    """builtins/src/poseidon/periodic.rs:242-290 (8-point NTT)."""
    coeffs = load_coeffs()
    with open(os.path.join(GOLDEN, "poseidon_round_keys.json")) as f:
        keys = {k: [[int(x, 16) for x in row] for row in v] for k, v in json.load(f).items()}
    for idx in range(3):
        first = [r[idx] for r in keys["FULL_ROUND_KEYS_1ST_HALF"]]
        second = [r[idx] for r in keys["FULL_ROUND_KEYS_2ND_HALF"]]
        first = first[1:] + [0]
        second = second[1:] + [0]
        evals = oracle.from_mont(oracle.ntt(oracle.to_mont(coeffs[f"FULL_ROUND_KEY_{idx}_COEFFS"])[None])[0])
        assert evals == first + second


def test_lde_is_coset_evaluation(oracle):
    rng = np.random.default_rng(3)
    cols = oracle.random_felts(rng, 1, 16)
    lde = oracle.from_mont(oracle.lde(cols, 1)[0])
    coeffs = oracle.from_mont(oracle.ntt(cols, inverse=True)[0])
    assert lde == ecref.naive_dft(coeffs + [0] * 16, offset=3)
    # the LDE restricted to ... the trace is recovered by evaluating on the non-coset domain
    assert ecref.naive_dft(coeffs) == oracle.from_mont(cols[0])


# ------------------------------------------------------------------- hash KATs
def test_blake2s_and_sha256_match_hashlib(oracle):
    rng = np.random.default_rng(4)
    for n in (0, 1, 31, 32, 63, 64, 65, 96, 127, 128, 224, 256, 288, 1000):
        data = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert oracle.hash_bytes(oracle.HASH_BLAKE2S, data) == hashlib.blake2s(data).digest()
        assert oracle.hash_bytes(oracle.HASH_SHA256, data) == hashlib.sha256(data).digest()
        masked = oracle.hash_bytes(oracle.HASH_BLAKE2S_M20, data)
        assert masked == b"\0" * 12 + hashlib.blake2s(data).digest()[12:]


def test_keccak_known_answers(oracle):
    # Keccak-256("") and ("abc"): public test vectors of the pre-NIST padding
    assert oracle.hash_bytes(oracle.HASH_KECCAK, b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert oracle.hash_bytes(oracle.HASH_KECCAK, b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    d = oracle.hash_bytes(oracle.HASH_KECCAK, b"x" * 300)
    assert oracle.hash_bytes(oracle.HASH_KECCAK_M20, b"x" * 300) == d[:20] + b"\0" * 12


def test_solidity_public_coin_draw_kat(oracle):
    """crypto/src/public_coin/solidity.rs:173-192: zero seed -> four draws.
    draw = Keccak(digest || BE32(counter)), reject >= 31p, interpret as Montgomery limbs (:88-97)."""
    expected = [
        914053382091189896561965228399096618375831658573140010954888220151670628653,
        3496720894051083870907112578962849417100085660158534559258626637026506475074,
        1568281537905787801632546124130153362941104398120976544423901633300198530772,
        539395842685339476048032152056539303790683868668644006005689195830492067187,
    ]
    digest, counter, got = b"\0" * 32, 0, []
    while len(got) < 4:
        raw = oracle.hash_bytes(oracle.HASH_KECCAK, digest + counter.to_bytes(32, "big"))
        counter += 1
        v = int.from_bytes(raw, "big")
        if v < 31 * P:
            got.append((v % P) * oracle.R_INV % P)
    assert got == expected


def test_cairo_public_coin_reseed_kat(oracle):
    """crypto/src/public_coin/cairo.rs:190-208."""
    seed = bytes([0x1f, 0x9c, 0x7b, 0xc9, 0xad, 0x41, 0xb8, 0xa6, 0x92, 0x36, 0x00, 0x6e, 0x7e, 0xea, 0x80, 0x38,
                  0xae, 0xa4, 0x32, 0x96, 0x07, 0x41, 0xb8, 0x19, 0x79, 0x16, 0x36, 0xf8, 0x2c, 0xc2, 0xd2, 0x5d])
    element = 941210603170996043151108091873286171552595656949
    data = (int.from_bytes(seed, "big") + 1).to_bytes(32, "big") + element.to_bytes(32, "big")
    want = bytes([0x60, 0x57, 0x79, 0xf6, 0xc9, 0xae, 0x87, 0x1e, 0xd7, 0x30, 0x56, 0xb4, 0xeb, 0xaa, 0x61, 0xa7,
                  0x7e, 0x7f, 0xb5, 0x09, 0xbc, 0x08, 0xc1, 0x93, 0xf1, 0x3a, 0xdc, 0xbf, 0x0c, 0x0b, 0xed, 0xc0])
    assert oracle.hash_bytes(oracle.HASH_BLAKE2S, data) == want


def test_pedersen_hash_kat(oracle):
    """builtins/src/pedersen/mod.rs:184-211 (StarkWare signature_test_data.json vectors)."""
    kats = [
        (1740729136829561885683894917751815192814966525555656371386868611731128807883,
         919869093895560023824014392670608914007817594969197822578496829435657368346,
         1382171651951541052082654537810074813456022260470662576358627909045455537762),
        (2514830971251288745316508723959465399194546626755475650431255835704887319877,
         3405079826265633459083097571806844574925613129801245865843963067353416465931,
         2962565761002374879415469392216379291665599807391815720833106117558254791559),
    ]
    for a, b, h in kats:
        assert oracle.pedersen_hash(a, b) == h
        assert ecref.pedersen_hash(a, b) == h
    rng = np.random.default_rng(5)
    for _ in range(5):
        a, b = (int.from_bytes(rng.bytes(32), "big") % P for _ in range(2))
        assert oracle.pedersen_hash(a, b) == ecref.pedersen_hash(a, b)
    for a, b in ((0, 0), (0, 1), (P - 1, P - 1), (2**248, 2**248 - 1)):
        assert oracle.pedersen_hash(a, b) == ecref.pedersen_hash(a, b)


# ----------------------------------------------------------------- Merkle trees
def _be32(v):
    return v.to_bytes(32, "big")


def _mont_int(row):
    return sum(int(row[i]) << (64 * i) for i in range(4))


@pytest.mark.parametrize("kind,n_cols", [("keccak", 2), ("keccak", 1), ("keccak_m20", 7), ("keccak_m20", 1)])
def test_leaf_variant_tree_structure(oracle, kind, n_cols):
    """Mirror of crypto/src/merkle/mod.rs:456-634 (8-row trees; round-trip there, structure here):
    recompute every node independently in Python and compare."""
    tree_kind = {"keccak": oracle.TREE_KECCAK, "keccak_m20": oracle.TREE_KECCAK_M20}[kind]
    mask = (lambda d: d[:20] + b"\0" * 12) if kind == "keccak_m20" else (lambda d: d)
    H = lambda data: mask(oracle.hash_bytes(oracle.HASH_KECCAK, data))
    rng = np.random.default_rng(6)
    cols = oracle.random_felts(rng, n_cols, 8)
    nodes, leaves, root = oracle.merkle_build(tree_kind, cols)
    if n_cols == 1:
        lv = [_be32(_mont_int(cols[0, i])) for i in range(8)]
        level = [H(lv[2 * i] + lv[2 * i + 1]) for i in range(4)]
    else:
        lv = [H(b"".join(_be32(_mont_int(cols[j, i])) for j in range(n_cols))) for i in range(8)]
        assert [bytes(l) for l in leaves] == lv
        level = [H(lv[2 * i] + lv[2 * i + 1]) for i in range(4)]
    assert [bytes(nodes[4 + i]) for i in range(4)] == level
    level2 = [H(level[0] + level[1]), H(level[2] + level[3])]
    assert [bytes(nodes[2]), bytes(nodes[3])] == level2
    assert bytes(nodes[1]) == H(level2[0] + level2[1]) == root


@pytest.mark.parametrize("n_friendly", [0, 1, 2, 3])
def test_friendly_tree_structure(oracle, n_friendly):
    """crypto/src/merkle/mod.rs:505-634 use N_FRIENDLY in {0,1,2,3} on 8-row matrices."""
    rng = np.random.default_rng(7)
    cols = oracle.random_felts(rng, 2, 8)
    nodes, leaves, root = oracle.merkle_build(oracle.TREE_FRIENDLY, cols, n_friendly=n_friendly)
    b2 = lambda data: b"\0" * 12 + hashlib.blake2s(data).digest()[12:]
    lv = [b2(_be32(_mont_int(cols[0, i])) + _be32(_mont_int(cols[1, i]))) for i in range(8)]
    assert [bytes(l) for l in leaves] == lv
    # depth 2 (built from leaves), depth 1, depth 0 (root); algebraic iff depth < n_friendly
    cur, cur_alg = lv, False
    got_levels = {2: [nodes[4 + i] for i in range(4)], 1: [nodes[2], nodes[3]], 0: [nodes[1]]}
    for depth in (2, 1, 0):
        alg = depth < n_friendly
        nxt = []
        for i in range(len(cur) // 2):
            a, b = cur[2 * i], cur[2 * i + 1]
            if not alg:
                nxt.append(b2(a + b))
            else:
                if not cur_alg:
                    a, b = int.from_bytes(a, "big"), int.from_bytes(b, "big")
                nxt.append(ecref.pedersen_hash(a, b))
        for g, w in zip(got_levels[depth], nxt):
            if alg:
                assert oracle.from_mont(np.frombuffer(bytes(g), dtype=np.uint64))[0] == w
            else:
                assert bytes(g) == w
        cur, cur_alg = nxt, alg
    assert root == (cur[0].to_bytes(32, "big") if cur_alg else cur[0])


def test_friendly_single_column_is_all_pedersen(oracle):
    rng = np.random.default_rng(8)
    cols = oracle.random_felts(rng, 1, 4)
    nodes, leaves, root = oracle.merkle_build(oracle.TREE_FRIENDLY, cols, n_friendly=22)
    v = oracle.from_mont(cols[0])
    he = lambda a, b: ecref.pedersen_hash(ecref.pedersen_hash(ecref.pedersen_hash(0, a), b), 2)   # hash/pedersen.rs:67-76
    l0, l1 = he(v[0], v[1]), he(v[2], v[3])
    assert oracle.from_mont(np.frombuffer(bytes(nodes[2]), dtype=np.uint64))[0] == l0
    assert root == ecref.pedersen_hash(l0, l1).to_bytes(32, "big")


def test_bitrev_rows_option(oracle):
    rng = np.random.default_rng(9)
    cols = oracle.random_felts(rng, 3, 16)
    perm = [int(f"{i:04b}"[::-1], 2) for i in range(16)]
    _, leaves_br, root_br = oracle.merkle_build(oracle.TREE_KECCAK_M20, cols, bitrev_rows=True)
    _, leaves_nat, root_nat = oracle.merkle_build(oracle.TREE_KECCAK_M20, np.ascontiguousarray(cols[:, perm]))
    assert np.array_equal(leaves_br, leaves_nat) and root_br == root_nat
