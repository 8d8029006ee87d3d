"""Host mirror of the reference's public coins (sandstorm_b200/public_coin.py) against the reference's own known-answer
tests, and of the host hashes (sandstorm_b200/hostcrypto.py) against public vectors and the C oracle."""
import os
import random

from sandstorm_b200 import hostcrypto as hc
from sandstorm_b200.public_coin import CairoVerifierPublicCoin, SolidityVerifierPublicCoin, leading_zeros, public_input_elements

P = hc.P


def test_host_hashes_known_answers(oracle):
    assert hc.keccak256(b"").hex() == "c5d2460186f7233c927e7db2dcc703c0e500b653ca82273b7bfad8045d85a470"
    assert hc.keccak256(b"abc").hex() == "4e03657aea45a94fc7d47ba826c8d667c0d1e6e33a64a036ec44f58fa12d6c45"
    rnd = random.Random(1)
    for n in (0, 1, 135, 136, 137, 271, 272, 300, 1000):
        data = bytes(rnd.randrange(256) for _ in range(n))
        assert hc.keccak256(data) == oracle.hash_bytes(oracle.HASH_KECCAK, data)
        assert hc.blake2s(data) == oracle.hash_bytes(oracle.HASH_BLAKE2S, data)
    # builtins/src/pedersen/mod.rs:184-211 (StarkWare signature_test_data.json)
    a = 0x03d937c035c878245caf64531a5756109c53068da139362728feb561405371cb
    b = 0x0208a0a10250e382e1e4bbe2880906c2791bf6275695e02fbbc6aeff9cd8b31a
    assert hc.pedersen_hash(a, b) == 0x030e480bed5fe53fa909cc0f8c4d99b8f9f2c016be4c41e13a4848797979c662
    a = 0x58f580910a6ca59b28927c08fe6c43e2e303ca384badc365795fc645d479d45
    b = 0x78734f65a067be9bdb39de18434d71e79f7b6466a4b66bbd979ab9e7515fe0b
    assert hc.pedersen_hash(a, b) == 0x68cc0b76cddd1dd4ed2301ada9b7c872b23875d5ff837b3a87993e0d9996b87


def test_solidity_coin_kat():
    """crypto/src/public_coin/solidity.rs:173-192."""
    coin = SolidityVerifierPublicCoin(b"\0" * 32)
    assert [coin.draw() for _ in range(4)] == [
        914053382091189896561965228399096618375831658573140010954888220151670628653,
        3496720894051083870907112578962849417100085660158534559258626637026506475074,
        1568281537905787801632546124130153362941104398120976544423901633300198530772,
        539395842685339476048032152056539303790683868668644006005689195830492067187]


def test_cairo_coin_kat():
    """crypto/src/public_coin/cairo.rs:190-208."""
    coin = CairoVerifierPublicCoin(bytes.fromhex("1f9c7bc9ad41b8a69236006e7eea8038aea432960741b819791636f82cc2d25d"))
    coin.reseed_with_bytes((941210603170996043151108091873286171552595656949).to_bytes(32, "big"))
    assert coin.digest.hex() == "605779f6c9ae871ed73056b4ebaa61a77e7fb509bc08c193f13adcbf0c0bedc0" and coin.counter == 0


def test_queries_and_proof_of_work_semantics():
    coin = CairoVerifierPublicCoin(b"\x11" * 32)
    q = coin.draw_queries(65, 1 << 19)
    assert q == sorted(set(q)) and len(q) <= 65 and all(0 <= v < 1 << 19 for v in q)
    assert coin.counter == (68 * 8 + 31) // 32                     # 68 = 65 rounded up to a multiple of 4
    coin = SolidityVerifierPublicCoin(b"\x22" * 32)
    nonce = next(k for k in range(1, 1 << 20) if coin.verify_proof_of_work(8, k))
    assert leading_zeros(hc.keccak256(coin._pow_prefix(8) + nonce.to_bytes(8, "big"))) >= 8
    assert not any(coin.verify_proof_of_work(8, k) for k in range(1, nonce))
    assert leading_zeros(b"\x00\x10" + b"\xff" * 30) == 11


def test_seed_from_the_example_public_input():
    """src/lib.rs:156-166 on the reference's example: the element list has the documented shape (src/input.rs) and the
    coin built from it is deterministic."""
    from oracle import cairo

    pi = cairo.AirPublicInput.from_file(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "array_sum", "air-public-input.json"))
    el = public_input_elements(pi, 12345)
    assert el[:4] == [14, 32764, 32770, 2110234636557836973669] and el[4:14] == [1, 5, 45, 76, 76, 76, 76, 76, 460, 460]
    assert el[14:] == [2508, 2508, 1, 0x40780017fff7fff, 1, 44, 12345]
    a, b = CairoVerifierPublicCoin.from_public_input(pi), CairoVerifierPublicCoin.from_public_input(pi)
    assert a.digest == b.digest and a.draw() == b.draw()
