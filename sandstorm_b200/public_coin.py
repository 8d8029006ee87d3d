"""Fiat-Shamir public coins of the reference (host side; sequential, a few hundred hashes per proof):

    SolidityVerifierPublicCoin   crypto/src/public_coin/solidity.rs:21-161   (Keccak-256; EthVerifier claims)
    CairoVerifierPublicCoin      crypto/src/public_coin/cairo.rs:27-174      (Blake2s-256 + Pedersen; CairoVerifier claims)

with the `ministark::random::PublicCoin` method names (new / reseed_with_digest / reseed_with_field_elements /
reseed_with_field_element_vector / reseed_with_int / draw / draw_queries / grind_proof_of_work /
verify_proof_of_work) and the seeding from the public input (src/lib.rs:145-167, src/input.rs:7-150).
Field elements cross this interface as canonical integers.  Both classes are pinned by the reference's own known-answer
tests (solidity.rs:173-192, cairo.rs:190-208: tests/test_public_coin.py).  The proof-of-work search is the GPU kernel
(`ss_pow_grind`), which returns the smallest nonce — the sequential build's answer (solidity.rs:137-138)."""
from __future__ import annotations

from . import hostcrypto as hc

P = hc.P
R = 2**256
_RINV = pow(R, -1, P)
POW_PREFIX = 0x0123456789ABCDED


def leading_zeros(digest: bytes) -> int:
    """ministark::random::leading_zeros: zero bits from the most significant bit of byte 0."""
    v = int.from_bytes(digest, "big")
    return 8 * len(digest) - v.bit_length()


class _Coin:
    HASH = None           # bytes -> 32 bytes
    POW_KIND = None       # ss_pow_grind hash_kind

    def __init__(self, digest: bytes):
        assert len(digest) == 32
        self.digest, self.counter = bytes(digest), 0

    # reseed_with_bytes / draw_bytes (solidity.rs:36-51, cairo.rs:42-57)
    def reseed_with_bytes(self, data: bytes) -> None:
        d = (int.from_bytes(self.digest, "big") + 1) % R
        self.digest, self.counter = self.HASH(d.to_bytes(32, "big") + bytes(data)), 0

    def draw_bytes(self) -> bytes:
        out = self.HASH(self.digest + self.counter.to_bytes(32, "big"))
        self.counter += 1
        return out

    def reseed_with_digest(self, digest: bytes) -> None:
        self.reseed_with_bytes(digest)

    def reseed_with_field_element_vector(self, vals) -> None:
        self.reseed_with_bytes(b"".join(hc.felt_bytes(v) for v in vals))

    def reseed_with_int(self, val: int) -> None:
        self.reseed_with_bytes(int(val).to_bytes(8, "big"))

    def draw(self) -> int:
        """rejection-sample below 31 p, then read the integer as Montgomery limbs (from_montgomery, utils.rs:9-12)."""
        while True:
            v = int.from_bytes(self.draw_bytes(), "big")
            if v < 31 * P:
                return v % P * _RINV % P

    def _ints(self, count: int):
        out, buf = [], b""
        while len(out) < count:
            if len(buf) < 8:
                buf += self.draw_bytes()
            out.append(int.from_bytes(buf[:8], "big"))
            buf = buf[8:]
        return out

    def _pow_prefix(self, bits: int) -> bytes:
        return self.HASH(POW_PREFIX.to_bytes(8, "big") + self.digest + bytes([bits]))

    def verify_proof_of_work(self, bits: int, nonce: int) -> bool:
        return leading_zeros(self.HASH(self._pow_prefix(bits) + int(nonce).to_bytes(8, "big"))) >= bits

    def grind_proof_of_work(self, bits: int, ctx=None) -> int:
        from .pow import grind_proof_of_work

        return grind_proof_of_work(self.digest, bits, self.POW_KIND, ctx)


class SolidityVerifierPublicCoin(_Coin):
    HASH = staticmethod(hc.keccak256)
    POW_KIND = 0

    def reseed_with_field_elements(self, vals) -> None:
        for v in vals:                                         # one reseed per element (solidity.rs:64-69)
            self.reseed_with_bytes(hc.felt_bytes(v))

    def draw_queries(self, max_n: int, domain_size: int) -> list[int]:
        return sorted({v % domain_size for v in self._ints(max_n)})        # BTreeSet (solidity.rs:99-118)

    @classmethod
    def from_public_input(cls, public_input) -> "SolidityVerifierPublicCoin":
        """src/lib.rs:145-154: Keccak over the BE32 public-input elements, main page hashed with CanonicalKeccak256HashFn
        (canonical, not Montgomery, element bytes: crypto/src/hash/keccak.rs:124-133)."""
        page = hc.keccak256(b"".join(int(v % P).to_bytes(32, "big") for e in public_input.public_memory for v in e))
        elements = public_input_elements(public_input, int.from_bytes(page, "big"))
        return cls(hc.keccak256(b"".join(e.to_bytes(32, "big") for e in elements)))


class CairoVerifierPublicCoin(_Coin):
    HASH = staticmethod(hc.blake2s)
    POW_KIND = 1

    def reseed_with_field_elements(self, vals) -> None:
        """Pedersen hash chain of the elements, then one reseed with its canonical BE bytes (cairo.rs:77-81)."""
        self.reseed_with_bytes(hc.pedersen_hash_elements(vals).to_bytes(32, "big"))

    def draw_queries(self, max_n: int, domain_size: int) -> list[int]:
        """the Cairo verifier samples in batches of 4 and truncates (cairo.rs:108-131)."""
        ints = self._ints((max_n + 3) // 4 * 4)[:max_n]
        return sorted({v % domain_size for v in ints})

    @classmethod
    def from_public_input(cls, public_input) -> "CairoVerifierPublicCoin":
        """src/lib.rs:156-166: Blake2s over the BE32 public-input elements, main page hashed with PedersenHashFn."""
        page = hc.pedersen_hash_elements(v for e in public_input.public_memory for v in e)
        elements = public_input_elements(public_input, page)
        return cls(hc.blake2s(b"".join(e.to_bytes(32, "big") for e in elements)))


SHARP_LAYOUT_CODE = {"starknet": 8319381555716711796, "recursive": 2110234636557836973669}       # binary/src/lib.rs:93-96


def public_input_elements(pi, main_page_hash: int) -> list[int]:
    """CairoAuxInput::public_input_elements (src/input.rs:10-150).  pi: an object with the fields of AirPublicInput
    (n_steps, rc_min, rc_max, layout, memory_segments {name: (begin_addr, stop_ptr)}, public_memory [(addr, value)])."""
    seg = pi.memory_segments
    vals = [pi.n_steps.bit_length() - 1, pi.rc_min, pi.rc_max, SHARP_LAYOUT_CODE[pi.layout]]
    for name in ("program", "execution", "output", "pedersen", "range_check"):
        vals += list(seg[name])
    padding = next(e for e in pi.public_memory if e[0] == 1)
    if pi.layout == "starknet":
        for name in ("ecdsa", "bitwise", "ec_op", "poseidon"):
            vals += list(seg[name])
    elif pi.layout == "recursive":
        vals += list(seg["bitwise"])
    else:
        raise NotImplementedError(pi.layout)
    vals += [padding[0], padding[1] % P, 1]                      # padding address, value, number of memory pages
    vals += [len(pi.public_memory), main_page_hash]              # main page: size, hash (no address: implicitly 1)
    return vals
