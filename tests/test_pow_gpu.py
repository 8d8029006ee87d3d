"""GPU: proof-of-work grinding against a brute-force restatement of the reference's definition
(crypto/src/public_coin/solidity.rs:120-160, cairo.rs:133-170) built on the oracle's hashes."""
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


def _leading_zero_bits(h: bytes) -> int:
    v = int.from_bytes(h, "big")
    return 256 - v.bit_length()


def _is_valid(oracle, kind, digest, bits, nonce):
    H = lambda b: oracle.hash_bytes(oracle.HASH_KECCAK if kind == 0 else oracle.HASH_BLAKE2S, b)
    prefix = H(bytes.fromhex("0123456789ABCDED") + digest + bytes([bits]))
    return _leading_zero_bits(H(prefix + nonce.to_bytes(8, "big"))) >= bits


@pytest.mark.parametrize("kind", [0, 1])
@pytest.mark.parametrize("bits", [0, 1, 8, 12])
def test_smallest_nonce_matches_brute_force(ss, oracle, kind, bits):
    from sandstorm_b200.pow import grind_proof_of_work

    rng = np.random.default_rng(100 * kind + bits)
    for _ in range(3):
        digest = rng.bytes(32)
        got = grind_proof_of_work(digest, bits, kind)
        want = next(n for n in range(1, 1 << 20) if _is_valid(oracle, kind, digest, bits, n))
        assert got == want


@pytest.mark.parametrize("kind", [0, 1])
def test_production_difficulty(ss, oracle, kind):
    """16 bits is the CLI default (cli/src/main.rs:55-56), 22 bits what crypto/benches/public_coin.rs grinds."""
    from sandstorm_b200.pow import grind_proof_of_work

    digest = bytes(range(32))
    for bits in (16, 22):
        nonce = grind_proof_of_work(digest, bits, kind)
        assert nonce >= 1 and _is_valid(oracle, kind, digest, bits, nonce)
