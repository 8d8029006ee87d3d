"""`Proof` container and its wire format: what `sandstorm prove` writes with
`proof.serialize_compressed` (cli/src/main.rs:206-207) and `sandstorm verify` reads (cli/src/main.rs:174).

ministark's `Proof` is not vendored; the layout below was recovered from — and is pinned by — the two proofs the
reference ships (bootloader-proof.bin, example/array-sum.proof.saved; copies under tests/golden/reference_proofs/):
`Proof.deserialize` consumes both to the last byte and `serialize` reproduces them byte for byte
(tests/test_reference_proof.py).  Encoding = ark-serialize compressed: integers little-endian, `Vec<T>` = u64 length +
items, field elements = 32-byte little-endian CANONICAL integers (not Montgomery), `SerdeOutput` digests = u64 32 + 32 bytes.

    options               5 x u8: num_queries, lde_blowup_factor, grinding_factor, fri_folding_factor, fri_max_remainder_coeffs
                          (ProofOptions::new argument order, cli/src/main.rs:152-158)
    trace_len             u64
    base_trace_commitment, Option<extension_trace_commitment> (u8 tag), composition_trace_commitment      digests
    fri_proof             layers: Vec<{ flattened_rows: Vec<Fq>, proofs: Vec<MerkleProof>, commitment: Digest }>,
                          remainder_coeffs: Vec<Fq>
    pow_nonce             u64
    trace_queries         base / extension / composition row values (Vec<F> each, query-major), then their Vec<MerkleProof>
    execution_trace_ood_evals, composition_trace_ood_evals                                                  Vec<Fq>

    MerkleProof           u8 variant (crypto/src/merkle/mod.rs:175-238, 354-417: 0 = hashed rows / MultiCol, 1 = raw single-column
                          leaves) | path: Vec<Digest>, LEAF LEVEL FIRST | sibling leaf | leaf        (leaf = digest, or Fq when raw)
    Digest                LeafVariant trees: SerdeOutput.  Friendly trees: MixedMerkleDigest = u8 tag (0 = HighLevel Pedersen felt,
                          1 = LowLevel Blake2s SerdeOutput; crypto/src/merkle/mixed.rs:46-71); single-column Friendly proofs use
                          bare Pedersen felts.

One Merkle proof per query (no multiproof), positions in ascending order (BTreeSet); FRI layer l lists the distinct
positions >> (3 l) — see sandstorm_b200/verify.py for what the positions mean."""
from __future__ import annotations

import struct
from dataclasses import dataclass, field

P = 2**251 + 17 * 2**192 + 1

HASHED, UNHASHED = 0, 1


@dataclass
class MerkleProof:
    variant: int                    # HASHED (row digests) | UNHASHED (raw single-column leaves)
    path: list                      # sibling digests, leaf level first: bytes (byte digest) or int (Pedersen felt)
    sibling: object                 # bytes digest / int felt (UNHASHED)
    leaf: object


@dataclass
class FriLayerProof:
    flattened_rows: list            # canonical ints, fold-many per opened row
    proofs: list
    commitment: object


@dataclass
class Proof:
    num_queries: int
    lde_blowup_factor: int
    grinding_factor: int
    fri_folding_factor: int
    fri_max_remainder_coeffs: int
    trace_len: int
    base_root: object
    ext_root: object                # None when the AIR has no extension columns
    comp_root: object
    fri_layers: list = field(default_factory=list)
    remainder_coeffs: list = field(default_factory=list)
    pow_nonce: int = 0
    base_values: list = field(default_factory=list)
    ext_values: list = field(default_factory=list)
    comp_values: list = field(default_factory=list)
    base_proofs: list = field(default_factory=list)
    ext_proofs: list = field(default_factory=list)
    comp_proofs: list = field(default_factory=list)
    ood_trace: list = field(default_factory=list)
    ood_comp: list = field(default_factory=list)
    friendly: bool = False          # digest codec: FriendlyMerkleTree (Cairo-verifier claims) vs LeafVariantMerkleTree

    # ---- writer -------------------------------------------------------------------------------------------
    def serialize(self) -> bytes:
        out = bytearray()
        u64 = lambda v: out.extend(struct.pack("<Q", v))
        felt = lambda v: out.extend(int(v % P).to_bytes(32, "little"))

        def digest(d, single_col=False):
            if not self.friendly:
                u64(32); out.extend(d)
            elif single_col:                     # MerkleView<PedersenDigest, Fp>
                felt(d)
            elif isinstance(d, int):             # MixedMerkleDigest::HighLevel
                out.append(0); felt(d)
            else:
                out.append(1); u64(32); out.extend(d)

        def felts(vs):
            u64(len(vs))
            for v in vs:
                felt(v)

        def proofs(ps):
            u64(len(ps))
            for p in ps:
                out.append(p.variant)
                u64(len(p.path))
                for d in p.path:
                    digest(d, p.variant == UNHASHED)
                for leaf in (p.sibling, p.leaf):
                    if p.variant == UNHASHED:
                        felt(leaf)
                    elif self.friendly:          # MultiCol leaves are plain Blake2s digests (SerdeOutput)
                        u64(32); out.extend(leaf)
                    else:
                        digest(leaf)

        out.extend(bytes([self.num_queries, self.lde_blowup_factor, self.grinding_factor, self.fri_folding_factor, self.fri_max_remainder_coeffs]))
        u64(self.trace_len)
        digest(self.base_root)
        if self.ext_root is None:
            out.append(0)
        else:
            out.append(1); digest(self.ext_root)
        digest(self.comp_root)
        u64(len(self.fri_layers))
        for layer in self.fri_layers:
            felts(layer.flattened_rows)
            proofs(layer.proofs)
            digest(layer.commitment)
        felts(self.remainder_coeffs)
        u64(self.pow_nonce)
        felts(self.base_values); felts(self.ext_values); felts(self.comp_values)
        proofs(self.base_proofs); proofs(self.ext_proofs); proofs(self.comp_proofs)
        felts(self.ood_trace); felts(self.ood_comp)
        return bytes(out)

    # ---- reader --------------------------------------------------------------------------------------------
    @classmethod
    def deserialize(cls, data: bytes, friendly: bool = False) -> "Proof":
        o = 0

        def u64():
            nonlocal o
            v = struct.unpack_from("<Q", data, o)[0]
            o += 8
            return v

        def felt():
            nonlocal o
            v = int.from_bytes(data[o:o + 32], "little")
            o += 32
            if v >= P:
                raise ValueError("non-canonical field element")
            return v

        def raw32():
            nonlocal o
            if u64() != 32:
                raise ValueError("digest length")
            d = data[o:o + 32]
            o += 32
            return d

        def digest(single_col=False):
            nonlocal o
            if not friendly:
                return raw32()
            if single_col:
                return felt()
            tag = data[o]
            o += 1
            if tag == 0:
                return felt()
            if tag == 1:
                return raw32()
            raise ValueError("MixedMerkleDigest tag")

        def felts():
            return [felt() for _ in range(u64())]

        def proofs():
            nonlocal o
            out = []
            for _ in range(u64()):
                variant = data[o]
                o += 1
                if variant not in (HASHED, UNHASHED):
                    raise ValueError("MerkleProof variant")
                path = [digest(variant == UNHASHED) for _ in range(u64())]
                leaves = [felt() if variant == UNHASHED else (raw32() if friendly else digest()) for _ in range(2)]
                out.append(MerkleProof(variant, path, leaves[0], leaves[1]))
            return out

        opts = list(data[:5])
        o = 5
        trace_len = u64()
        base_root = digest()
        tag = data[o]
        o += 1
        ext_root = digest() if tag else None
        comp_root = digest()
        layers = []
        for _ in range(u64()):
            rows = felts()
            ps = proofs()
            layers.append(FriLayerProof(rows, ps, digest()))
        remainder = felts()
        nonce = u64()
        bv, ev, cv = felts(), felts(), felts()
        bp, ep, cp = proofs(), proofs(), proofs()
        ood_t, ood_c = felts(), felts()
        if o != len(data):
            raise ValueError(f"{len(data) - o} trailing bytes")
        return cls(*opts, trace_len, base_root, ext_root, comp_root, layers, remainder, nonce, bv, ev, cv, bp, ep, cp, ood_t, ood_c, friendly)


# ---- assembling a Proof from the prover's result (sandstorm_b200/prover.py HotPathResult with keep_openings=True) ---------------
_R = 2**256
_RINV = pow(_R, -1, P)


def _felts(arr) -> list:
    """uint64[..., 4] Montgomery limbs -> canonical ints (flattened, row-major)."""
    a = arr.reshape(-1, 4)
    return [(int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) * _RINV % P for r in a]


def _wire_digest(raw: bytes, algebraic: bool):
    """storage form (csrc/merkle.cu: byte digest, or Montgomery limbs of a Pedersen felt) -> wire form (bytes / canonical int)"""
    return int.from_bytes(raw, "little") * _RINV % P if algebraic else bytes(raw)


def _merkle_proofs(kind, n_friendly, rows, paths, row_digest) -> list:
    """rows: uint64[q, n_cols, 4]; paths: uint8[q, depth, 32] as ss_merkle_open returns them (sibling leaf, then sibling nodes)."""
    from . import _lib

    friendly = kind == _lib.TREE_FRIENDLY
    q, n_cols = rows.shape[0], rows.shape[1]
    height = paths.shape[1]
    out = []
    for k in range(q):
        if n_cols == 1:
            leaf, sibling = _felts(rows[k])[0], int.from_bytes(bytes(paths[k, 0]), "little") * _RINV % P
            path = [_wire_digest(bytes(paths[k, j]), friendly) for j in range(1, height)]
            out.append(MerkleProof(UNHASHED, path, sibling, leaf))
        else:
            path = [_wire_digest(bytes(paths[k, j]), friendly and height - j < n_friendly) for j in range(1, height)]
            out.append(MerkleProof(HASHED, path, bytes(paths[k, 0]), row_digest(kind, _felts(rows[k]))))
    return out


def assemble_proof(res, options, trace_len: int) -> Proof:
    """options: prover.ProofOptions; res: HotPathResult of prove(..., keep_openings=True).  Roots of Friendly trees travel as
    MixedMerkleDigest::HighLevel felts (their `as_bytes` is the big-endian canonical integer ss_merkle_root returns)."""
    from . import _lib
    from .verify import row_digest

    kind, nf = options.tree_kind, options.n_friendly
    friendly = kind == _lib.TREE_FRIENDLY

    def root(b: bytes, log_rows: int, n_cols: int):
        return int.from_bytes(b, "big") if friendly and (n_cols == 1 or nf > 0) else b

    log_N = (trace_len << options.log_blowup).bit_length() - 1
    tq = res.trace_queries
    layers = []
    for k, lay in enumerate(res.fri_layers):
        layers.append(FriLayerProof(_felts(lay["rows"]), _merkle_proofs(kind, nf, lay["rows"], lay["paths"], row_digest),
                                    root(res.fri_roots[k], 0, 1 << options.log_fold)))
    return Proof(options.num_queries, 1 << options.log_blowup, options.grinding_factor, 1 << options.log_fold, options.max_remainder_coeffs, trace_len,
                 root(res.roots["base"], log_N, tq["base"]["rows"].shape[1]),
                 root(res.roots["ext"], log_N, tq["ext"]["rows"].shape[1]) if "ext" in res.roots else None,
                 root(res.roots["composition"], log_N, 2), layers, _felts(res.remainder), res.pow_nonce,
                 _felts(tq["base"]["rows"]), _felts(tq["ext"]["rows"]), _felts(tq["composition"]["rows"]),
                 _merkle_proofs(kind, nf, tq["base"]["rows"], tq["base"]["paths"], row_digest),
                 _merkle_proofs(kind, nf, tq["ext"]["rows"], tq["ext"]["paths"], row_digest),
                 _merkle_proofs(kind, nf, tq["composition"]["rows"], tq["composition"]["paths"], row_digest),
                 list(res.ood_trace), list(res.ood_composition), friendly)
