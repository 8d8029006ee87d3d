"""Host-side hash functions of the reference's `crypto/src/hash` — used where the reference itself works on the
CPU on a handful of bytes: the Fiat-Shamir public coins (crypto/src/public_coin), Merkle-path verification
(`MerkleTree::verify`, crypto/src/merkle/mod.rs:125-165, 306-346) and the public-input seed (src/lib.rs:145-167).
The bulk hashing of the hot path (leaves, tree levels, proof-of-work search) is CUDA (csrc/hashes.cuh, merkle.cu,
pow.cu); nothing here is on a timed path.

    keccak256      sha3 0.10.8 `Keccak256` (legacy 0x01 padding, rate 136)            crypto/src/hash/keccak.rs:10-48
    blake2s        blake2 0.10.6 `Blake2s256`                                        crypto/src/hash/blake2s.rs:1-48
    pedersen_hash  starknet-crypto 0.6.1 `pedersen_hash`, points P0..P4              builtins/src/pedersen/mod.rs:31-36, constants.rs:6-29
    masks          hash/mod.rs:5-23 (Keccak keeps bytes 0..19, Blake2s keeps bytes 12..31)
"""
from __future__ import annotations

import hashlib

P = 2**251 + 17 * 2**192 + 1
R = 2**256

_RC = [0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
       0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
       0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
       0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008]
_ROT = [[0, 36, 3, 41, 18], [1, 44, 10, 45, 2], [62, 6, 43, 15, 61], [28, 55, 25, 21, 56], [27, 20, 39, 8, 14]]
_M64 = (1 << 64) - 1


def _keccak_f(a):
    rol = lambda v, n: ((v << n) | (v >> (64 - n))) & _M64 if n else v
    for rc in _RC:
        c = [a[x][0] ^ a[x][1] ^ a[x][2] ^ a[x][3] ^ a[x][4] for x in range(5)]
        d = [c[(x - 1) % 5] ^ rol(c[(x + 1) % 5], 1) for x in range(5)]
        a = [[a[x][y] ^ d[x] for y in range(5)] for x in range(5)]
        b = [[0] * 5 for _ in range(5)]
        for x in range(5):
            for y in range(5):
                b[y][(2 * x + 3 * y) % 5] = rol(a[x][y], _ROT[x][y])
        a = [[b[x][y] ^ (~b[(x + 1) % 5][y] & b[(x + 2) % 5][y]) for y in range(5)] for x in range(5)]
        a[0][0] ^= rc
    return a


def keccak256(data: bytes) -> bytes:
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    msg.extend(b"\0" * (-len(msg) % rate))
    msg[-1] |= 0x80
    a = [[0] * 5 for _ in range(5)]
    for off in range(0, len(msg), rate):
        for k in range(rate // 8):
            a[k % 5][k // 5] ^= int.from_bytes(msg[off + 8 * k:off + 8 * k + 8], "little")
        a = _keccak_f(a)
    return b"".join(a[k % 5][k // 5].to_bytes(8, "little") for k in range(4))


def blake2s(data: bytes) -> bytes:
    return hashlib.blake2s(data, digest_size=32).digest()


def sha256(data: bytes) -> bytes:
    return hashlib.sha256(data).digest()


def mask_keccak20(d: bytes) -> bytes:
    """mask_least_significant_bytes::<20> (hash/mod.rs:5-13): bytes 0..19 kept."""
    return d[:20] + b"\0" * 12


def mask_blake20(d: bytes) -> bytes:
    """mask_most_significant_bytes::<20> (hash/mod.rs:15-23): bytes 12..31 kept."""
    return b"\0" * 12 + d[12:]


def felt_bytes(v: int) -> bytes:
    """hash_elements / coin encoding of a field element: 32-byte big-endian of the MONTGOMERY limbs
    (to_montgomery(v).to_be_bytes, crypto/src/utils.rs:15-17)."""
    return (v % P * R % P).to_bytes(32, "big")


# ---- Pedersen hash on the StarkWare curve y^2 = x^3 + x + b ----------------------------------------------------------
_POINTS = [
    (2089986280348253421170679821480865132823066470938446095505822317253594081284, 1713931329540660377023406109199410414810705867260802078187082345529207694986),
    (996781205833008774514500082376783249102396023663454813447423147977397232763, 1668503676786377725805489344771023921079126552019160156920634619255970485781),
    (2251563274489750535117886426533222435294046428347329203627021249169616184184, 1798716007562728905295480679789526322175868328062420237419143593021674992973),
    (2138414695194151160943305727036575959195309218611738193261179310511854807447, 113410276730064486255102093846540133784865286929052426931474106396135072156),
    (2379962749567351885752724891227938183011949129833673362440656643086021394946, 776496453633298175483985398648758586525933812536653089401905292063708816422),
]


def _ec_add(p1, p2):
    if p1 is None:
        return p2
    if p2 is None:
        return p1
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1 + 1) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


_chains: list = []


def _doubling_chains():
    if not _chains:
        for base, bits in ((_POINTS[1], 248), (_POINTS[2], 4), (_POINTS[3], 248), (_POINTS[4], 4)):
            pts, acc = [], base
            for _ in range(bits):
                pts.append(acc)
                acc = _ec_add(acc, acc)
            _chains.append(pts)
    return _chains


def pedersen_hash(a: int, b: int) -> int:
    """x-coordinate of P0 + a_lo*P1 + a_hi*P2 + b_lo*P3 + b_hi*P4 (lo = low 248 bits, hi = top 4 bits)."""
    ch = _doubling_chains()
    acc = _POINTS[0]
    for k, v in enumerate((a % P, b % P)):
        lo, hi = v & (2**248 - 1), v >> 248
        for i in range(248):
            if (lo >> i) & 1:
                acc = _ec_add(acc, ch[2 * k][i])
        for i in range(4):
            if (hi >> i) & 1:
                acc = _ec_add(acc, ch[2 * k + 1][i])
    return acc[0]


def pedersen_hash_elements(elements) -> int:
    """PedersenHashFn::hash_elements (crypto/src/hash/pedersen.rs:67-76): h = H(h, e) from 0, then H(h, count)."""
    h, k = 0, 0
    for e in elements:
        h = pedersen_hash(h, e)
        k += 1
    return pedersen_hash(h, k)
