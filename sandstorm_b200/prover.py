"""GPU hot path of `Stark::prove` — the stages of SURVEY.md §3.1 that run on the device, in the order
ministark's prover runs them (steps 3-5, 8-10, 11-13, 15), for a layout of the reference
(plain / recursive / starknet) and the proof options of the CLI (cli/src/main.rs:51-60):

    base trace  : LDE -> Merkle commit                      [challenges drawn]
    ext trace   : LDE -> Merkle commit                      [composition coefficient drawn]
    composition : constraint evaluation over the LDE coset -> coset iNTT -> ce interleaved columns
                  -> LDE -> Merkle commit                   [OOD point z drawn]
    OOD         : trace polynomials at z*g^offset for every tap, composition columns at z^ce
    DEEP        : sum alpha^i (T(x) - y) / (x - z g^k) on the sub-coset 3<w_n>, extended to the LDE coset   (src/lib.rs:102-116)
    FRI         : per layer commit (rows of `fold` evaluations) -> alpha -> fold, until
                  layer_size / blowup <= max_remainder
    queries     : Merkle openings + rows at the query positions

The transcript is driven by a public coin with the reference's `PublicCoin` interface (sandstorm_b200/public_coin.py:
the Solidity- and Cairo-verifier coins; `SeededCoin` below is a stand-in that draws from a seeded generator for
benchmarks on synthetic columns).  The order of reseeds and draws follows ministark's `ProverChannel` as recalled
(ministark is not vendored; DESIGN.md §2 lists it as unpinned):

    reseed(base root) -> challenges -> [extension columns built from them] -> reseed(ext root) -> composition coefficient
    -> reseed(composition root) -> z -> reseed_with_field_elements(trace OOD values), (composition OOD values) -> DEEP alpha
    -> per FRI layer: reseed(layer root), draw fold alpha -> reseed_with_field_element_vector(remainder)
    -> proof-of-work nonce, reseed_with_int(nonce) -> draw_queries

The extension trace may be given as a callable(challenges) -> Matrix (`Trace::build_extension_columns`, on the device:
sandstorm_b200/ext_columns.py) and the hints as a callable(challenges) -> list (`AirConfig::gen_hints`).  The composition
program is compiled ahead of time as a template (air/program.py ProgramTemplate); only its value patch runs between the
extension commitment and constraint evaluation."""
from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .air import compile_program, compile_template
from .air.program import tap_reach
from .air.deep import DEEP_FILTER_MIN_TAPS, deep_expr_filtered, deep_expr_shifted, deep_expr_symbolic, deep_filter_columns, deep_terms
from .air.evaluate import evaluate
from .air.expr import P
from .air.layouts import load_layout
from .matrix import Matrix, fri_fold, inv_x_minus_c, ood_eval, poly_eval

R = 2**256


def _brev(v: int, bits: int) -> int:
    return int(f"{v:0{bits}b}"[::-1], 2) if bits else 0


def _mont(v: int) -> np.ndarray:
    m = v % P * R % P
    return np.array([(m >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


@dataclass
class ProofOptions:
    """cli/src/main.rs:51-60 defaults."""
    num_queries: int = 65
    log_blowup: int = 1
    log_fold: int = 3
    max_remainder_coeffs: int = 16
    grinding_factor: int = 0                     # cli default 16 (cli/src/main.rs:55); 0 skips the search
    ce_blowup: int = 2                           # composition columns (src/lib.rs:110-113 air.ce_blowup_factor())
    capi_collectives: bool = False               # world > 1: LDE / halo / commit / all-gather through the C ABI's own NCCL calls
                                                 # (csrc/dist.cu: what a Rust host uses) instead of torch.distributed
    tree_kind: int = _lib.TREE_KECCAK_M20        # src/claims.rs:18-21 (starknet / EthVerifier); recursive claims use TREE_FRIENDLY
    n_friendly: int = 22                         # NUM_FRIENDLY_COMMITMENT_LAYERS, src/claims.rs:10 (TREE_FRIENDLY only)
    col_pad_rows: int = 0                        # padding between the columns of the working matrix (not a protocol parameter)
    deep_filter_min_taps: int = DEEP_FILTER_MIN_TAPS   # DEEP quotient: a column with at least this many mask offsets gets its sum of
                                                 # poles W_c(x) = sum_t a_t / (x - z g^off_t) — and the layout its V(x) when it has that many
                                                 # distinct offsets — from two size-n transforms instead of one multiply-add per pole per row
                                                 # (air/deep.py deep_expr_filtered; same values; 0 disables)
    ood_transform_min_taps: int = 32             # a column with at least this many mask offsets gets its out-of-domain values from ONE
                                                 # transform onto the coset z<g> (ss_coset_eval) instead of one n-term sum per offset;
                                                 # same values either way (0 disables)


class SeededCoin:
    """Stand-in public coin for synthetic benchmarks: draws from a seeded generator (the same on every rank), ignores
    reseeds, skips the proof-of-work search."""

    def __init__(self, seed: int = 0xB200):
        self.rnd = random.Random(seed)

    def draw(self) -> int:
        return self.rnd.randrange(P)

    def reseed_with_digest(self, digest): pass
    def reseed_with_field_elements(self, vals): pass
    def reseed_with_field_element_vector(self, vals): pass
    def reseed_with_int(self, val): pass
    def grind_proof_of_work(self, bits, ctx=None): return 0

    def draw_queries(self, max_n: int, domain_size: int) -> list[int]:
        return sorted({self.rnd.randrange(domain_size) for _ in range(max_n)})


@dataclass
class HotPathResult:
    roots: dict = field(default_factory=dict)
    fri_roots: list = field(default_factory=list)
    remainder: np.ndarray | None = None
    ood_trace: list = field(default_factory=list)
    ood_composition: list = field(default_factory=list)
    query_positions: list = field(default_factory=list)
    opened_bytes: int = 0
    deep_matches_full_evaluation: bool | None = None
    composition_top_zero: bool | None = None
    challenges: list = field(default_factory=list)
    hints: list = field(default_factory=list)
    composition_coeffs: list = field(default_factory=list)
    ood_point: int = 0
    deep_alpha: int = 0
    fri_alphas: list = field(default_factory=list)
    pow_nonce: int = 0
    # openings at the query positions (the `Queries` + `FriProof` payload of ministark's Proof)
    trace_queries: dict = field(default_factory=dict)      # name -> {"rows": uint64[q, cols, 4], "paths": uint8[q, depth, 32]}
    fri_layers: list = field(default_factory=list)          # per layer {"positions", "rows": uint64[q, fold, 4], "paths"}


class HotPathProver:
    def __init__(self, layout: str, log_n: int, options: ProofOptions | None = None, seed: int = 0xB200, device=None,
                 rank: int = 0, world: int = 1, coin=None):
        self.layout = load_layout(layout)
        self.log_n, self.opt = log_n, options or ProofOptions()
        self.n, self.N = 1 << log_n, 1 << (log_n + self.opt.log_blowup)
        self.ce = self.opt.ce_blowup                            # air.ce_blowup_factor(): 2 for Cairo's degree-2 constraints
        if (1 << self.opt.log_blowup) < self.ce:
            raise ValueError("the LDE blowup must be at least the constraint-evaluation blowup")
        self.coin = coin if coin is not None else SeededCoin(seed)      # same seed on every rank: identical challenges everywhere
        self.rank, self.world = rank, world
        if world & (world - 1):
            raise ValueError("world size must be a power of two")
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.g = pow(3, (P - 1) // self.n, P)
        # columns of the working matrix: trace | composition (ce) | w = 1/(x-1) | u = 1/(x-z) | v = 1/(x-z^ce)
        C = self.layout.num_columns
        self.comp_col, self.w_col, self.u_col, self.v_col = C, C + self.ce, C + self.ce + 1, C + self.ce + 2
        # ... | V | W_c of the tap-heavy columns   (only when the DEEP quotient takes its pole sums from transforms)
        self.n_work_cols = C + self.ce + 3
        self.value_col, self.filter_cols = None, {}
        taps = self.layout.taps()
        mt = self.opt.deep_filter_min_taps
        if mt and log_n >= 1 and (deep_filter_columns(taps, mt) or len({off for _, off in taps}) >= mt):
            self.value_col = self.n_work_cols
            self.filter_cols = {col: self.value_col + 1 + j for j, col in enumerate(deep_filter_columns(taps, mt))}
            self.n_work_cols += 1 + len(self.filter_cols)
        self._deep_template_direct = None
        self._template = self._deep_template = None
        self._composition_program = None
        self._challenges = self._hints = self._alpha = None
        self.timeline: list = []

    # ---- what this rank reads of the trace (for callers that stream it from the host) -----------------------------
    def trace_rows_needed(self) -> list[tuple[int, int]]:
        """(first row, count) ranges of the trace rows this rank reads of EVERY column: its block-cyclic pieces
        (parallel.pieces) plus, after each piece, the max_offset rows the out-of-domain dot products reach (wrapping mod n).
        For callers that stream the trace from host memory (bench.py e2e leg)."""
        if self.world == 1:
            return [(0, self.n)]
        from .parallel import pieces

        out, reach = [], self.layout.max_offset
        for lo, cnt in pieces(self.log_n, self.rank, self.world):
            hi = lo + cnt + reach
            out.append((lo, min(hi, self.n) - lo))
            if hi > self.n:
                out.append((0, min(hi - self.n, self.n)))
        return out

    # ---- host-side stand-ins for the public coin -------------------------------------------------------
    def _draw(self) -> int:
        return self.coin.draw()

    def mark(self, name: str):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.timeline.append((name, ev))

    def composition_template(self):
        """AirConfig::constraints + composition_constraint compiled with the challenges, hints and the composition
        coefficient left open: depends on (layout, n, blowup) only, so it is built ahead of the proof."""
        if self._template is None:
            L = self.layout
            # boundary denominators X - g^e read the auxiliary column w = 1/(x - 1) (see Layout.constraints)
            self._template = compile_template(L.composition(self.n, inv_x_minus_one_col=self.w_col), self.log_n, self.opt.log_blowup,
                                              L.n_challenges(), L.n_hints(), 1)
        return self._template

    def deep_template(self):
        """the DEEP quotient program (src/lib.rs:102-116 coefficients) with alpha and the out-of-domain values left open:
        like the composition, only a value patch is left for the proof itself."""
        if self._deep_template is None:
            L = self.layout
            taps = L.taps()
            if self.value_col is None:
                expr = deep_expr_symbolic(taps, self.ce, self.comp_col, self.u_col, self.v_col, self.g, P)
            else:
                expr = deep_expr_filtered(taps, self.ce, self.comp_col, self.u_col, self.v_col, self.g, P, self.filter_cols, self.value_col)
            self._deep_template = compile_template(expr, self.log_n, self.opt.log_blowup, 1, len(taps) + self.ce, 1)
        return self._deep_template

    def deep_template_direct(self):
        """the per-row form of the same quotient (every pole a shifted read of u): what the self-check compares against"""
        if self.value_col is None:
            return self.deep_template()
        if self._deep_template_direct is None:
            taps = self.layout.taps()
            self._deep_template_direct = compile_template(deep_expr_symbolic(taps, self.ce, self.comp_col, self.u_col, self.v_col, self.g, P),
                                                          self.log_n, self.opt.log_blowup, 1, len(taps) + self.ce, 1)
        return self._deep_template_direct

    def _pole_sums_on_coset(self, jobs, z: int) -> None:
        """jobs: [(weights {offset: w}, out)];  out[i] = sum_off weights[off] / (x_i - z g^off) on the n points x_i = 3 g^i
        (out: a [n, 4] view, any stride; with world > 1 only the owned pieces are written).
        Every x_i has x_i^n = K = 3^n and every pole zeta = z g^off has zeta^n = z^n, so on the coset
            1 / (x - zeta) = C sum_{k<n} x^k zeta^(n-1-k),   C = 1 / (K - z^n),   zeta^(n-1-k) = z^(n-1-k) g^-off g^(-off k)
        and the sum is the polynomial  C z^(n-1) sum_k (3/z)^k B[k] g^(i k)  with  B[k] = sum_off (weights[off] g^-off) g^(-off k):
        the unnormalised inverse transform of a sparse vector, a geometric scaling and a forward transform — the local LDE
        (ss_ntt_shard stages 1 + 2 fused) with c0 = C z^(n-1), h0 = 3/z and no expansion, whatever the number of poles.
        world > 1: the same pair split over the ranks (parallel.ShardedTransforms)."""
        from .parallel import DeviceShardOps, pieces

        n, log_n, g, dev = self.n, self.log_n, self.g, self.device
        K, zn = pow(3, n, P), pow(z, n, P)
        c0 = pow((K - zn) % P, -1, P) * pow(z, n - 1, P) % P
        h0 = 3 * pow(z, -1, P) % P
        W, r = self.world, self.rank
        mine = pieces(log_n, r, W) if W > 1 else [(0, n)]
        cache = self.__dict__.setdefault("_ginv_pow", {})          # g^-off per mask offset: layout constants, computed once

        def ginv(off):
            v = cache.get(off)
            if v is None:
                v = cache[off] = pow(g, -off, P)
            return v
        from .parallel import ShardedTransforms

        buf = torch.empty((n, 4), dtype=torch.int64, device=dev)
        res = torch.empty((n, 4), dtype=torch.int64, device=dev) if W == 1 else None
        if W > 1 and getattr(self, "_pole_st", None) is None:
            # (software-pipelining these pairs over the two LDE pipes was measured SLOWER at 8 GPUs — 57 vs 30 ms for four
            #  pairs — so they run back to back on the caller's stream)
            self._pole_st = ShardedTransforms(r, W, DeviceShardOps(self.ctx), dev)
        for weights, out in jobs:
            acc: dict[int, int] = {}
            for off, w in weights.items():
                acc[off % n] = (acc.get(off % n, 0) + w * ginv(off)) % P
            idx = torch.tensor(sorted(acc), dtype=torch.int64, device=dev)
            vals = torch.from_numpy(np.stack([_mont(acc[o]) for o in sorted(acc)]).view(np.int64)).to(dev)
            for lo, cnt in mine:
                buf[lo:lo + cnt].zero_()
            buf[idx] = vals                                  # (entries outside the owned pieces are never read)
            if W == 1:
                DeviceShardOps(self.ctx).ntt_shard(buf, log_n, 3, 0, c0, h0, None, res)
                out.copy_(res)
                continue
            st = self._pole_st
            share = st.to_coefficients(buf, log_n, c0 * pow(h0, r, P) % P, pow(h0, W, P))
            st.from_coefficients(share, log_n - (W.bit_length() - 1), 0, buf)
            for lo, cnt in mine:
                out[lo:lo + cnt].copy_(buf[lo:lo + cnt])

    def prepare(self):
        """everything that depends only on (layout, trace length, options): call once, ahead of the proofs."""
        self.composition_template()
        self.deep_template()

    def composition_program(self, challenges, hints, alpha):
        """the per-proof value patch (a few milliseconds): constants <- challenges, hints, composition coefficient."""
        self._composition_program = self.composition_template().patch(challenges, hints, alpha)
        return self._composition_program

    # ---- commitment helper: whole tree on one GPU; on several, row-range sub-trees + combined root ----------------------
    def _commit(self, ptr: int, col_stride: int, n_cols: int, log_rows: int, order: int = _lib.ORDER_BITREV):
        """Returns (root bytes, handle).  ptr: device address of column 0, row 0 of the matrix.  With world > 1 this is used
        for the FRI layers only, whose evaluations every rank holds completely (they are all-gathered), so both the sharded
        and the small unsharded build read valid rows on every rank."""
        c, opt = self.ctx, self.opt
        world, rank = self.world, self.rank
        shard = world > 1 and log_rows - (world.bit_length() - 1) >= 10
        rows = 1 << log_rows
        lo, cnt = (rank * (rows // world), rows // world) if shard else (0, rows)
        handle = ctypes.c_void_p()
        # Pedersen levels are counted from the root of the WHOLE tree: a sub-tree below log2(world) levels keeps the rest
        friendly = 0
        if opt.tree_kind == _lib.TREE_FRIENDLY:
            friendly = max(0, opt.n_friendly - (world.bit_length() - 1)) if shard else opt.n_friendly
        if shard:
            # this rank's range of TREE leaves (leaf p = row brev(p)): digests, then the sub-tree over them
            leaves = torch.empty((cnt, 4), dtype=torch.int64, device=self.device)
            c.check(c.lib.ss_hash_rows(c.handle, opt.tree_kind, ctypes.c_void_p(ptr), col_stride, n_cols, log_rows, order, lo, cnt,
                                       ctypes.c_void_p(leaves.data_ptr()), None))
            c.check(c.lib.ss_merkle_build_from_leaves(c.handle, opt.tree_kind, friendly, ctypes.c_void_p(leaves.data_ptr()), cnt.bit_length() - 1,
                                                      ctypes.byref(handle), None))
        else:
            c.check(c.lib.ss_merkle_build(c.handle, opt.tree_kind, friendly, ctypes.c_void_p(ptr), col_stride, n_cols, log_rows, order,
                                          ctypes.byref(handle), None))
        root = (ctypes.c_uint8 * 32)()
        c.check(c.lib.ss_merkle_root(c.handle, handle, root))
        if shard:
            from .parallel import gather_subroots

            subs = gather_subroots(bytes(root), world, self.device)
            buf = (ctypes.c_uint8 * (32 * world)).from_buffer_copy(b"".join(subs))
            c.check(c.lib.ss_merkle_combine(c.handle, opt.tree_kind, buf, world.bit_length() - 1, root))
        self._last_subroots = b"".join(subs) if shard else None         # (None: this rank holds the whole tree)
        return bytes(root), handle

    def _remainder(self, evals: torch.Tensor, log_blowup: int) -> np.ndarray:
        """last FRI layer (natural order on offset<w_m>) -> the coefficients of f(offset * X), degree < m / blowup."""
        m = evals.shape[0]
        coeffs = Matrix(evals.clone().view(1, m, 4), self.ctx).ntt_(inverse=True).data[0]
        keep = max(1, m >> log_blowup)
        self.remainder_high_zero = not bool(coeffs[keep:].any().item())       # a low-degree codeword: the dropped half vanishes
        return coeffs[:keep].cpu().numpy().view(np.uint64)

    def _gather_rows(self, full: torch.Tensor, lo: int, cnt: int):
        """all-gather of a row-sharded vector: every rank contributes full[lo:lo+cnt]."""
        if self.world == 1:
            return
        import torch.distributed as dist

        dist.all_gather_into_tensor(full, full[lo:lo + cnt].clone())

    # ---- the device stages ---------------------------------------------------------------------------------
    def prove(self, base: Matrix, ext, queries: bool = True, self_check: bool = False, column_ready=None, hints=None,
              keep_openings: bool = False) -> HotPathResult:
        """base: the base trace columns; ext: the extension columns, or a callable(challenges) -> Matrix that builds them
        once the challenges are drawn (Trace::build_extension_columns); hints: list or callable(challenges) -> list
        (AirConfig::gen_hints; None = drawn from the coin, for synthetic columns).  keep_openings: return the opened rows
        and authentication paths (the proof payload) instead of only counting their bytes.
        column_ready(k): optional hook called before trace column k (base then extension; None = all) is first read, so that a
        caller streaming the trace from host memory can order its uploads against the LDE (bench.py e2e leg).
        With world > 1 the work is row-sharded over the ranks (_prove_sharded)."""
        opt, L = self.opt, self.layout
        assert base.num_cols == L.num_base_columns and base.num_rows == self.n
        if self.world > 1:
            return self._prove_sharded(base, ext, hints, column_ready, queries, keep_openings)
        coin = self.coin
        res = HotPathResult()
        dev = self.device
        n, N, b = self.n, self.N, opt.log_blowup
        c = self.ctx = base.ctx
        nb, C = L.num_base_columns, L.num_columns
        handles = []
        self.mark("start")
        # one matrix for every committed column: trace | composition (ce) | w | u | v  (see __init__)
        S = N + opt.col_pad_rows                         # column stride of the working matrix
        all_lde = torch.empty((self.n_work_cols, S, 4), dtype=torch.int64, device=dev)[:, :N]
        lde = all_lde[:C]

        def lde_cols(src: Matrix, first_col: int):
            for j in range(src.num_cols):
                if column_ready is not None:
                    column_ready(first_col + j)          # e.g. make the stream wait for the upload of this column
                keep = None
                if first_col + j in heavy:               # interpolated polynomial (c_k 3^k, bit-reversed) kept for the OOD stage
                    keep = coeffs_of[first_col + j] = torch.empty((n, 4), dtype=torch.int64, device=dev)
                c.check(c.lib.ss_lde(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src.data[j].data_ptr()), src.col_stride, 1, self.log_n, b,
                                     ctypes.c_void_p(lde[first_col + j].data_ptr()), S, ctypes.c_void_p(keep.data_ptr()) if keep is not None else None, n,
                                     _lib.ORDER_NATURAL, None))

        taps = L.taps()
        per_col = {}
        for col, _ in taps:
            per_col[col] = per_col.get(col, 0) + 1
        heavy = {col for col, cnt in per_col.items() if opt.ood_transform_min_taps and cnt >= opt.ood_transform_min_taps and self.log_n >= 1}
        coeffs_of = {}
        # 3-5: base trace
        lde_cols(base, 0)
        self.mark("lde_base")
        res.roots["base"], h = self._commit(lde.data_ptr(), S, nb, self.log_n + b); handles.append(h)
        self.mark("merkle_base")
        # 6-7: challenges, hints, extension columns
        coin.reseed_with_digest(res.roots["base"])
        challenges = res.challenges = [coin.draw() for _ in range(L.n_challenges())]
        if callable(ext):
            ext = ext(challenges)
        assert ext.num_cols == L.num_extension_columns and ext.num_rows == self.n
        if hints is None:
            hints = [coin.draw() for _ in range(L.n_hints())]
        elif callable(hints):
            hints = hints(challenges)
        res.hints = list(hints)
        self.mark("ext_columns")
        # 8: extension trace
        lde_cols(ext, nb)
        self.mark("lde_ext")
        res.roots["ext"], h = self._commit(lde[nb].data_ptr(), S, C - nb, self.log_n + b); handles.append(h)
        self.mark("merkle_ext")
        coin.reseed_with_digest(res.roots["ext"])
        # 9: constraint evaluation, boundary denominators from w = 1/(x - 1)
        res.composition_coeffs = [coin.draw()]
        prog = self.composition_program(challenges, res.hints, res.composition_coeffs)
        self.mark("patch")
        inv_x_minus_c(all_lde[self.w_col], _mont(1), c)
        comp_evals = torch.empty((N, 4), dtype=torch.int64, device=dev)
        self.mark("inv_w")
        evaluate(prog, Matrix(all_lde, c), b, out=comp_evals)
        self.mark("constraint_eval")
        # 10: composition polynomial -> ce columns (coefficients j, j+ce, ...) -> LDE -> commit
        work = Matrix(comp_evals.view(1, N, 4), c)
        work.ntt_(inverse=True, coset=True)
        self.mark("ntt_comp_inv")
        if self_check and N > self.ce * n:
            # a trace that satisfies the AIR gives a composition polynomial of degree < ce * n: the upper coefficients vanish
            res.composition_top_zero = not bool(comp_evals[self.ce * n:].any().item())
        comp_coeffs = comp_evals[:self.ce * n].view(n, self.ce, 4).permute(1, 0, 2).contiguous()        # [ce, n, 4] natural order
        comp_lde = all_lde[self.comp_col:self.comp_col + self.ce]
        comp_lde.zero_()
        comp_lde[:, :n] = comp_coeffs
        self.mark("comp_split")
        Matrix(comp_lde, c).ntt_(coset=True)
        self.mark("ntt_comp_fwd")
        res.roots["composition"], h = self._commit(comp_lde.data_ptr(), S, self.ce, self.log_n + b); handles.append(h)
        self.mark("merkle_comp")
        # 11: out-of-domain evaluations of every tap, straight from the trace (barycentric dot products with one shared
        #     weight vector, ss_ood_eval)
        coin.reseed_with_digest(res.roots["composition"])
        z = res.ood_point = coin.draw()
        if column_ready is not None:
            column_ready(None)                           # every trace column is read from here on
        parts = np.zeros((len(taps), 4), dtype=np.uint64)
        for mat, first, count in ((base, 0, nb), (ext, nb, C - nb)):
            idx = [k for k, (col, _) in enumerate(taps) if first <= col < first + count and col not in heavy]
            if idx:
                parts[idx] = ood_eval(mat, [taps[k][0] - first for k in idx], [taps[k][1] for k in idx], _mont(z))
        if heavy:
            # columns with many offsets: T(z g^j) for every j from one transform of the kept coefficients (h = z / 3 undoes the
            # coset scaling ss_lde left on them), then the mask offsets are read out of it
            on_z = torch.empty((n, 4), dtype=torch.int64, device=dev)
            h_scale = _mont(z * pow(3, -1, P) % P)
            for col in sorted(heavy):
                c.check(c.lib.ss_coset_eval(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(coeffs_of[col].data_ptr()), n, 1, self.log_n,
                                            h_scale.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(on_z.data_ptr()), n, None))
                idx = [k for k, (tc, _) in enumerate(taps) if tc == col]
                rows_at = np.array([taps[k][1] % n for k in idx], dtype=np.uint64)
                got = np.zeros((len(idx), 1, 4), dtype=np.uint64)
                c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(on_z.data_ptr()), n, 1, rows_at.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
                                             len(idx), got.ctypes.data_as(ctypes.c_void_p)))
                parts[idx] = got[:, 0]
            del on_z
            coeffs_of.clear()
        to_int = lambda a: [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]
        ood_m = to_int(parts)
        zc = pow(z, self.ce, P)
        ood_c = poly_eval(Matrix(comp_coeffs, c), list(range(self.ce)), np.stack([_mont(zc)] * self.ce), natural_order=True)
        self.mark("ood")
        rinv = pow(R, -1, P)
        res.ood_trace, res.ood_composition = [v * rinv % P for v in ood_m], [v * rinv % P for v in to_int(ood_c)]
        # 12: DEEP composition over the LDE coset (coefficients = powers of one alpha, src/lib.rs:102-116)
        coin.reseed_with_field_elements(res.ood_trace)
        coin.reseed_with_field_elements(res.ood_composition)
        alpha = res.deep_alpha = coin.draw()
        # (the sub-coset evaluation below reads u and v only at rows that are multiples of the blowup)
        inv_x_minus_c(all_lde[self.u_col], _mont(z), c, log_row_step=b)
        inv_x_minus_c(all_lde[self.v_col], _mont(zc), c, log_row_step=b)
        deep_prog = self.deep_template().patch([alpha], res.ood_trace + res.ood_composition, [0])
        del comp_coeffs, comp_evals, work
        if self.value_col is not None:
            # the long pole sums of the quotient as columns on the sub-coset rows (deep_expr_filtered): V, then W_c per heavy column
            step = 1 << b
            v_w: dict[int, int] = {}
            col_w: dict[int, dict] = {col: {} for col in self.filter_cols}
            a_k = 1
            for (col, off), y in zip(taps, res.ood_trace):
                v_w[off] = (v_w.get(off, 0) + a_k * y) % P
                if col in col_w:
                    col_w[col][off] = (col_w[col].get(off, 0) + a_k) % P
                a_k = a_k * alpha % P
            self._pole_sums_on_coset([(v_w, all_lde[self.value_col, ::step])] +
                                     [(col_w[col], all_lde[fcol, ::step]) for col, fcol in self.filter_cols.items()], z)
        # The quotient has degree < n - 1, so its n values on the sub-coset 3<w_n> — the LDE rows that are multiples of
        # the blowup — determine it: evaluate only those (1/blowup of the work), then extend like any other column
        # (coset iNTT of size n, zero padding, coset NTT of size N).  Same polynomial, hence the same N evaluations.
        deep = torch.empty((N, 4), dtype=torch.int64, device=dev)
        self.mark("deep_setup")
        evaluate(deep_prog, Matrix(all_lde, c), b, out=deep[:n], log_row_step=b)
        self.mark("deep")
        Matrix(deep[:n].view(1, n, 4), c).ntt_(inverse=True, coset=True)
        deep[n:].zero_()
        Matrix(deep.view(1, N, 4), c).ntt_(coset=True)
        self.mark("deep_lde")
        if self_check:
            # test hook: the extended quotient equals the row-by-row evaluation on the whole LDE coset
            inv_x_minus_c(all_lde[self.u_col], _mont(z), c)
            inv_x_minus_c(all_lde[self.v_col], _mont(zc), c)
            direct = self.deep_template_direct().patch([alpha], res.ood_trace + res.ood_composition, [0])
            res.deep_matches_full_evaluation = bool(torch.equal(deep, evaluate(direct, Matrix(all_lde, c), b)))
        # 13: FRI layers.  Conventions pinned by the reference's proof artefacts (sandstorm_b200/verify.py): a layer commits, in
        #     bit-reversed row order, rows of the `fold` evaluations that fold together (ORDER_BITREV_RC on the evaluation
        #     buffer viewed with col_stride = rows: no data movement); the fold has no 1/fold factor; the remainder is sent as
        #     the coefficients of f(offset * X).
        evals, log_size, offset = deep, self.log_n + b, 3
        layers = []
        while (1 << log_size) >> b > opt.max_remainder_coeffs and log_size > opt.log_fold:
            rows = 1 << (log_size - opt.log_fold)
            root, handle = self._commit(evals.data_ptr(), rows, 1 << opt.log_fold, log_size - opt.log_fold, _lib.ORDER_BITREV_RC)
            res.fri_roots.append(root)
            coin.reseed_with_digest(root)
            fri_alpha = coin.draw()
            res.fri_alphas.append(fri_alpha)
            nxt = fri_fold(evals, opt.log_fold, _mont(fri_alpha), _mont(offset), starkware_scale=True, ctx=c)
            layers.append((handle, evals, log_size))
            evals, log_size, offset = nxt, log_size - opt.log_fold, pow(offset, 1 << opt.log_fold, P)
        self.final_domain = (log_size, offset)
        res.remainder = self._remainder(evals, b)
        coin.reseed_with_field_element_vector([v * rinv % P for v in to_int(res.remainder)])
        self.mark("fri")
        # 14: proof of work (GPU search for the smallest nonce), 15: query positions
        if opt.grinding_factor:
            res.pow_nonce = coin.grind_proof_of_work(opt.grinding_factor, c)
            coin.reseed_with_int(res.pow_nonce)
        if queries:
            # a position is a LEAF index: leaf p of every trace tree commits the LDE row brev(p) (bit-reversed commitment)
            pos = coin.draw_queries(opt.num_queries, N)
            res.query_positions = pos
            log_N = self.log_n + b
            idx = np.array(pos, dtype=np.uint64)
            nat = np.array([_brev(p, log_N) for p in pos], dtype=np.uint64)
            for name, h, (first, ncols) in zip(("base", "ext", "composition"), handles, ((0, nb), (nb, C - nb), (self.comp_col, self.ce))):
                paths = np.zeros((len(idx), log_N, 32), dtype=np.uint8)
                c.check(c.lib.ss_merkle_open(c.handle, h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                             paths.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
                rows_out = np.zeros((len(idx), ncols, 4), dtype=np.uint64)
                c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(all_lde[first].data_ptr()), S, ncols,
                                             nat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(nat), rows_out.ctypes.data_as(ctypes.c_void_p)))
                res.opened_bytes += paths.nbytes + rows_out.nbytes
                if keep_openings:
                    res.trace_queries[name] = {"rows": rows_out, "paths": paths}
            F = 1 << opt.log_fold
            col_perm = [_brev(j, opt.log_fold) for j in range(F)]
            for handle, layer_evals, ls in layers:
                log_rows = ls - opt.log_fold
                rows = 1 << log_rows
                pos = sorted({p >> opt.log_fold for p in pos})               # leaf indices of this layer
                idx = np.array(pos, dtype=np.uint64)
                nat = np.array([_brev(r, log_rows) for r in pos], dtype=np.uint64)
                paths = np.zeros((len(idx), log_rows, 32), dtype=np.uint8)
                c.check(c.lib.ss_merkle_open(c.handle, handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                             paths.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
                res.opened_bytes += paths.nbytes
                if keep_openings:
                    rows_out = np.zeros((len(idx), F, 4), dtype=np.uint64)
                    c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(layer_evals.data_ptr()), rows, F,
                                                 nat.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(nat), rows_out.ctypes.data_as(ctypes.c_void_p)))
                    res.fri_layers.append({"positions": pos, "rows": np.ascontiguousarray(rows_out[:, col_perm]), "paths": paths})
            self.mark("queries")
        for handle, _, _ in layers:
            c.lib.ss_tree_free(handle)
        for h in handles:
            c.lib.ss_tree_free(h)
        return res

    # ---- the device stages over several GPUs: every transform row-sharded (parallel.ShardedTransforms) ---------------------
    def _commit_pieces(self, cols: torch.Tensor, col_stride: int, log_rows: int):
        """Commitment of a block-cyclic matrix in the reference's (bit-reversed) leaf order.  A rank owns contiguous row
        ranges, but tree leaf p commits row brev(p), so a rank's rows are spread over the whole tree: every rank hashes the
        rows it owns, the 32-byte leaf digests are redistributed with one all-to-all (row i goes to the rank that owns leaf
        brev(i), i.e. brev_W(i mod W): 1/n_cols of the matrix traffic), each rank builds the sub-tree over its contiguous
        range of N / W leaves, and the W sub-roots are all-gathered and combined (ss_merkle_combine; Pedersen levels for
        the Friendly tree).  A single-column matrix is committed with raw leaves (the elements themselves travel).
        Returns (root, this rank's sub-tree handle, the W sub-roots): what _open_sharded needs for the query phase."""
        import torch.distributed as dist

        from .parallel import _all_to_all, pieces

        c, opt, W, r = self.ctx, self.opt, self.world, self.rank
        log_w = W.bit_length() - 1
        N = 1 << log_rows
        s = N // (W * W)
        n_cols, ptr = cols.shape[0], cols.data_ptr()             # cols: [n_cols, N, 4] view of the working matrix
        D = torch.empty((W, s, 4), dtype=torch.int64, device=self.device)             # my rows' leaves, natural order
        for k1, (lo, cnt) in enumerate(pieces(log_rows, r, W)):
            if n_cols == 1:
                D[k1].copy_(cols[0, lo:lo + cnt])
            else:
                c.check(c.lib.ss_hash_rows(c.handle, opt.tree_kind, ctypes.c_void_p(ptr), col_stride, n_cols, log_rows, _lib.ORDER_NATURAL, lo, cnt,
                                           ctypes.c_void_p(D[k1].data_ptr()), None))
        # row i = k1 m + r s + u W + c  ->  rank brev_W(c); as [k1][u][c] the c axis selects the destination
        by_c = D.view(W, s // W, W, 4)
        send = torch.stack([by_c[:, :, _brev(q, log_w)] for q in range(W)])              # [dest][k1][u]
        recv = torch.empty_like(send)                                                    # [src][k1][u]
        _all_to_all(recv.view(W, -1, 4), send.view(W, -1, 4), W)
        # j = i >> log W = k1 (m / W) + src (s / W) + u in natural order, then to tree order: local leaf = brev(j)
        leaves = recv.permute(1, 0, 2, 3).contiguous().view(N // W, 4)
        c.check(c.lib.ss_bitrev_permute32(c.handle, ctypes.c_void_p(leaves.data_ptr()), log_rows - log_w, None))
        friendly = max(0, opt.n_friendly - log_w) if opt.tree_kind == _lib.TREE_FRIENDLY else 0
        handle = ctypes.c_void_p()
        if n_cols == 1:
            c.check(c.lib.ss_merkle_build(c.handle, opt.tree_kind, friendly, ctypes.c_void_p(leaves.data_ptr()), N // W, 1, log_rows - log_w,
                                          _lib.ORDER_NATURAL, ctypes.byref(handle), None))
        else:
            c.check(c.lib.ss_merkle_build_from_leaves(c.handle, opt.tree_kind, friendly, ctypes.c_void_p(leaves.data_ptr()), log_rows - log_w,
                                                      ctypes.byref(handle), None))
        root = (ctypes.c_uint8 * 32)()
        c.check(c.lib.ss_merkle_root(c.handle, handle, root))
        from .parallel import gather_subroots

        subs = gather_subroots(bytes(root), W, self.device)
        buf = (ctypes.c_uint8 * (32 * W)).from_buffer_copy(b"".join(subs))
        c.check(c.lib.ss_merkle_combine(c.handle, opt.tree_kind, buf, log_w, root))
        return bytes(root), handle, b"".join(subs)

    def _open_sharded(self, handle, subroots, log_rows: int, n_cols: int, positions) -> np.ndarray:
        """MerkleTree::prove for leaf positions of a tree whose leaves are split into W contiguous ranges (one sub-tree per rank):
        the owner of a position opens its sub-tree and appends the siblings above its sub-root (ss_merkle_combine_open); the paths
        are then shared so that every rank holds uint8[q, log_rows, 32] exactly as ss_merkle_open on one GPU returns it.
        subroots None: every rank holds the whole tree (small FRI layers)."""
        import torch.distributed as dist

        c, opt, W, r = self.ctx, self.opt, self.world, self.rank
        u64p, u8p = ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint8)
        q = len(positions)
        paths = np.zeros((q, log_rows, 32), dtype=np.uint8)
        if subroots is None:
            idx = np.array(positions, dtype=np.uint64)
            c.check(c.lib.ss_merkle_open(c.handle, handle, idx.ctypes.data_as(u64p), q, paths.ctypes.data_as(u8p)))
            return paths
        log_w = W.bit_length() - 1
        sub_log = log_rows - log_w
        algebraic = int(opt.tree_kind == _lib.TREE_FRIENDLY and (n_cols == 1 or log_w < opt.n_friendly))
        buf = (ctypes.c_uint8 * len(subroots)).from_buffer_copy(subroots)
        if opt.capi_collectives:
            idx = np.array(positions, dtype=np.uint64)
            c.check(c.lib.ss_dist_open(c.handle, opt.tree_kind, algebraic, handle, buf, idx.ctypes.data_as(u64p), q, paths.ctypes.data_as(u8p), None))
            return paths
        mine = [k for k, p in enumerate(positions) if p >> sub_log == r]
        if mine:
            idx = np.array([positions[k] & ((1 << sub_log) - 1) for k in mine], dtype=np.uint64)
            low = np.zeros((len(mine), sub_log, 32), dtype=np.uint8)
            c.check(c.lib.ss_merkle_open(c.handle, handle, idx.ctypes.data_as(u64p), len(mine), low.ctypes.data_as(u8p)))
            top = np.zeros((log_w, 32), dtype=np.uint8)
            c.check(c.lib.ss_merkle_combine_open(c.handle, opt.tree_kind, buf, log_w, algebraic, r, top.ctypes.data_as(u8p)))
            paths[mine, :sub_log] = low
            paths[mine, sub_log:] = top
        t = torch.from_numpy(paths).to(self.device)
        dist.all_reduce(t)                                   # one contributor per entry: the sum is the value
        return t.cpu().numpy()

    def _read_rows_sharded(self, cols: torch.Tensor, col_stride: int, log_rows: int, rows) -> np.ndarray:
        """Matrix::read_row of a block-cyclic matrix: the rank that owns natural row i reads it, then the rows are shared.
        -> uint64[q, n_cols, 4] on every rank."""
        import torch.distributed as dist

        c, W, r = self.ctx, self.world, self.rank
        n_cols, q = cols.shape[0], len(rows)
        out = np.zeros((q, n_cols, 4), dtype=np.uint64)
        u64p = ctypes.POINTER(ctypes.c_uint64)
        if self.opt.capi_collectives:
            idx = np.array(rows, dtype=np.uint64)
            c.check(c.lib.ss_dist_gather_rows(c.handle, ctypes.c_void_p(cols.data_ptr()), col_stride, n_cols, log_rows, idx.ctypes.data_as(u64p), q,
                                              out.ctypes.data_as(ctypes.c_void_p), None))
            return out
        m = (1 << log_rows) // W
        s = m // W
        mine = [k for k, i in enumerate(rows) if (i % m) // s == r]
        if mine:
            idx = np.array([rows[k] for k in mine], dtype=np.uint64)
            got = np.zeros((len(mine), n_cols, 4), dtype=np.uint64)
            c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(cols.data_ptr()), col_stride, n_cols, idx.ctypes.data_as(u64p), len(mine),
                                         got.ctypes.data_as(ctypes.c_void_p)))
            out[mine] = got
        t = torch.from_numpy(out.view(np.int64)).to(self.device)
        dist.all_reduce(t)
        return t.cpu().numpy().view(np.uint64)

    def _lde_pipes(self, dev, W, rank):
        """two (stream, ShardedTransforms) pairs, each with its own context (scratch buffers and work-buffer pool are per context
        and recycled in stream order)."""
        if getattr(self, "_pipes", None) is None:
            from .context import Context
            from .parallel import DeviceShardOps, ShardedTransforms

            self._pipes = [(torch.cuda.Stream(), ShardedTransforms(rank, W, DeviceShardOps(Context(dev.index)), dev)) for _ in range(2)]
        return self._pipes

    def _gather_pieces(self, vec: torch.Tensor, log_len: int) -> None:
        """all-gather of a block-cyclic vector, in place: afterwards every rank holds every row."""
        import torch.distributed as dist

        from .parallel import pieces

        W = self.world
        ps = pieces(log_len, self.rank, W)
        s = ps[0][1]
        mine = torch.cat([vec[lo:lo + cnt] for lo, cnt in ps])                                   # [k1][s]
        every = torch.empty((W,) + tuple(mine.shape), dtype=vec.dtype, device=vec.device)
        dist.all_gather_into_tensor(every, mine)
        vec[:W * W * s].view(W, W, s, *vec.shape[1:]).copy_(every.view(W, W, s, *vec.shape[1:]).permute(1, 0, 2, *range(3, 2 + vec.dim())))

    def _exchange_halo(self, cols: torch.Tensor, log_len: int, halo: int) -> None:
        """cols: [n_cols, len, 4] block-cyclic.  After the call the `halo` rows that follow each owned piece (wrapping) are
        valid too: they are the first rows of the next rank's piece, received at their absolute position."""
        import torch.distributed as dist

        from .parallel import pieces

        if halo == 0:
            return
        W, r = self.world, self.rank
        s = pieces(log_len, r, W)[0][1]
        if halo > s:                                        # small domains: simply give every rank every row
            for j in range(cols.shape[0]):
                self._gather_pieces(cols[j], log_len)
            return
        prv, nxt = (r - 1) % W, (r + 1) % W
        ops, staged = [], []
        for j in range(cols.shape[0]):
            for (lo, _), (nlo, _) in zip(pieces(log_len, r, W), pieces(log_len, nxt, W)):
                ops.append(dist.P2POp(dist.isend, cols[j, lo:lo + halo], prv))
                ops.append(dist.P2POp(dist.irecv, cols[j, nlo:nlo + halo], nxt))
        for req in dist.batch_isend_irecv(ops):
            req.wait()
        del staged

    def _prove_sharded(self, base: Matrix, ext, hints, column_ready, queries: bool = True, keep_openings: bool = False) -> HotPathResult:
        """world > 1 (torch.distributed initialised, one process per GPU).  Every column's transforms are split over the
        ranks (parallel.ShardedTransforms: two all-to-alls per LDE column, 1/W of the arithmetic per rank whatever the
        number of columns); each rank ends up with W contiguous row ranges ("pieces") of EVERY column, on which it hashes
        leaves, evaluates the constraints and the DEEP quotient with no further exchange except a halo of
        max_offset * blowup rows per piece, the 32-byte sub-roots and the partial out-of-domain sums.  The FRI layers
        (1/8 of the work of one column and shrinking) run on the all-gathered DEEP evaluations.  Bit-identical to world = 1
        (tools/check_multi_gpu.py)."""
        import torch.distributed as dist

        from .parallel import DeviceShardOps, ShardedTransforms, pieces

        opt, L, coin = self.opt, self.layout, self.coin
        res = HotPathResult()
        dev, W, rank = self.device, self.world, self.rank
        n, N, b = self.n, self.N, opt.log_blowup
        log_n, log_N = self.log_n, self.log_n + b
        c = self.ctx = base.ctx
        nb, C = L.num_base_columns, L.num_columns
        if N // (W * W) < 16 << b:
            raise ValueError("trace too short to shard over this many GPUs")
        st = ShardedTransforms(rank, W, DeviceShardOps(c), dev)
        PN, Pn = pieces(log_N, rank, W), pieces(log_n, rank, W)
        halo = L.max_offset << b
        capi = opt.capi_collectives
        if capi and not getattr(c, "_dist_ready", False):
            ident = (ctypes.c_uint8 * 128)()
            if rank == 0:
                c.check(c.lib.ss_dist_unique_id(c.handle, ident))
            t = torch.tensor(list(ident), dtype=torch.uint8, device=dev)
            dist.broadcast(t, 0)                                            # (the id travels by the host's own means)
            ident = (ctypes.c_uint8 * 128)(*t.cpu().tolist())
            c.check(c.lib.ss_dist_init(c.handle, ident, rank, W))
            c._dist_ready = True
        from .matrix import _stream_ptr

        def lde_one(src_col, dst_col, on_coset=False):
            if capi:
                c.check(c.lib.ss_dist_lde(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src_col.data_ptr()), log_n, b, int(on_coset),
                                          ctypes.c_void_p(dst_col.data_ptr()), _stream_ptr()))
            else:
                st.lde(src_col, log_n, b, dst_col, src_on_coset=on_coset)

        def halo_of(cols):
            if capi and halo and halo <= PN[0][1]:
                c.check(c.lib.ss_dist_halo(c.handle, ctypes.c_void_p(cols.data_ptr()), S, cols.shape[0], log_N, halo, _stream_ptr()))
            else:
                self._exchange_halo(cols, log_N, halo)

        trace_trees = []                                  # (sub-tree handle, sub-roots, columns) of base / ext / composition

        def commit(cols):
            if capi:
                root, sub, subs = (ctypes.c_uint8 * 32)(), ctypes.c_void_p(), (ctypes.c_uint8 * (32 * W))()
                c.check(c.lib.ss_dist_commit(c.handle, opt.tree_kind, opt.n_friendly, ctypes.c_void_p(cols.data_ptr()), S, cols.shape[0], log_N,
                                             root, ctypes.byref(sub), subs, _stream_ptr()))
                root, subs = bytes(root), bytes(subs)
            else:
                root, sub, subs = self._commit_pieces(cols, S, log_N)
            trace_trees.append((sub, subs, cols))
            return root

        self.mark("start")
        S = N + opt.col_pad_rows
        all_lde = torch.empty((self.n_work_cols, S, 4), dtype=torch.int64, device=dev)[:, :N]
        lde = all_lde[:C]

        def lde_cols(src: Matrix, first_col: int):
            k = src.num_cols
            if capi or k == 1:
                for j in range(k):
                    if column_ready is not None:
                        column_ready(first_col + j)
                    lde_one(src.data[j], lde[first_col + j])
                return
            # software pipeline over the columns: while the local LDE of column j runs on one stream, the first exchange of
            # column j + 1 and the second exchange of column j - 1 proceed on the other (two contexts: separate scratch)
            main = torch.cuda.current_stream()
            pipes = self._lde_pipes(dev, W, rank)
            for stream, _ in pipes:
                stream.wait_stream(main)

            def begin(j):
                if column_ready is not None:
                    column_ready(first_col + j)
                stream, stx = pipes[j & 1]
                stream.wait_stream(main)                 # (column_ready made `main` wait for the column's upload)
                with torch.cuda.stream(stream):
                    stx.lde_begin(src.data[j], log_n, slot=j & 1)

            begin(0)
            for j in range(k):
                if j + 1 < k:
                    begin(j + 1)
                stream, stx = pipes[j & 1]
                with torch.cuda.stream(stream):
                    stx.lde_finish(log_n, b, lde[first_col + j], slot=j & 1)
            for stream, _ in pipes:
                main.wait_stream(stream)

        # 3-5: base trace
        lde_cols(base, 0)
        self.mark("lde_base")
        halo_of(lde[:nb])
        self.mark("share_base")
        res.roots["base"] = commit(lde[:nb])
        self.mark("merkle_base")
        coin.reseed_with_digest(res.roots["base"])
        challenges = res.challenges = [coin.draw() for _ in range(L.n_challenges())]
        if callable(ext):
            ext = ext(challenges)
        if hints is None:
            hints = [coin.draw() for _ in range(L.n_hints())]
        elif callable(hints):
            hints = hints(challenges)
        res.hints = list(hints)
        self.mark("ext_columns")
        # 8: extension trace
        lde_cols(ext, nb)
        self.mark("lde_ext")
        halo_of(lde[nb:C])
        self.mark("share_ext")
        res.roots["ext"] = commit(lde[nb:C])
        self.mark("merkle_ext")
        coin.reseed_with_digest(res.roots["ext"])
        # 9: constraint evaluation on the owned pieces
        res.composition_coeffs = [coin.draw()]
        prog = self.composition_program(challenges, res.hints, res.composition_coeffs)
        self.mark("patch")
        lo_w, hi_w = tap_reach(prog.blob, log_N).get(self.w_col, (0, 0))
        whole_w = hi_w - lo_w >= N // 4                   # tiny domains: signed offsets are ambiguous, take every row
        if whole_w:
            inv_x_minus_c(all_lde[self.w_col], _mont(1), c)
        comp_evals = torch.empty((N, 4), dtype=torch.int64, device=dev)
        for lo, cnt in PN:
            if not whole_w:
                inv_x_minus_c(all_lde[self.w_col], _mont(1), c, rows=(lo + lo_w, min(N, cnt + hi_w - lo_w)))
        self.mark("inv_w")
        for lo, cnt in PN:
            evaluate(prog, Matrix(all_lde, c), b, out=comp_evals, rows=(lo, cnt))
        self.mark("constraint_eval")
        # 10: composition polynomial -> ce interleaved columns -> LDE -> commit (all sharded)
        assert self.ce == 2
        comp_lde = all_lde[self.comp_col:self.comp_col + self.ce]
        shares = st.composition_columns(comp_evals, log_n, b, [comp_lde[0], comp_lde[1]])
        self.mark("ntt_comp")
        res.roots["composition"] = commit(comp_lde)
        self.mark("merkle_comp")
        # 11: out-of-domain values: partial barycentric sums over the owned trace rows, summed over the ranks
        coin.reseed_with_digest(res.roots["composition"])
        z = res.ood_point = coin.draw()
        taps = L.taps()
        if column_ready is not None:
            column_ready(None)
        to_int = lambda a: [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]
        rinv = pow(R, -1, P)
        acc = [0] * len(taps)
        per_col = {}
        for col, _ in taps:
            per_col[col] = per_col.get(col, 0) + 1
        # (a sharded pair costs two exchanges: it only pays for twice as many taps as on one GPU)
        heavy = {col for col, cnt in per_col.items() if opt.ood_transform_min_taps and cnt >= 2 * opt.ood_transform_min_taps}
        if heavy:
            # tap-heavy columns: T(z g^j) for every j by one more sharded transform pair (coefficients scaled by z^k, then the
            # forward transform without expansion); the mask offsets are then read from their owners.  Same values as the sums.
            on_z = torch.empty((n, 4), dtype=torch.int64, device=dev)
            ninv = pow(n, -1, P)
            for col in sorted(heavy):
                src = base.data[col] if col < nb else ext.data[col - nb]
                share = st.to_coefficients(src, log_n, ninv * pow(z, rank, P) % P, pow(z, W, P))
                st.from_coefficients(share, log_n - (W.bit_length() - 1), 0, on_z)
                idx = [k for k, (tc, _) in enumerate(taps) if tc == col]
                vals = to_int(self._read_rows_sharded(on_z.view(1, n, 4), n, log_n, [taps[k][1] % n for k in idx])[:, 0])
                if rank == 0:                             # (final values, not partial sums: counted once)
                    for k, v in zip(idx, vals):
                        acc[k] = v
            del on_z
        for mat, first, count in ((base, 0, nb), (ext, nb, C - nb)):
            idx = [k for k, (col, _) in enumerate(taps) if first <= col < first + count and col not in heavy]
            if not idx:
                continue
            for lo, cnt in Pn:
                part = to_int(ood_eval(mat, [taps[k][0] - first for k in idx], [taps[k][1] for k in idx], _mont(z), rows=(lo, cnt)))
                for k, v in zip(idx, part):
                    acc[k] = (acc[k] + v) % P
        # composition columns: each rank evaluates its residue class of the coefficients (coset-scaled, bit-reversed: the
        # ss_lde coefficient format) at y = z^ce:  sum_j2 x[r + W j2] (y/3)^(r + W j2)
        zc = pow(z, self.ce, P)
        y3 = zc * pow(3, -1, P) % P
        point = 3 * pow(y3, W, P) % P
        for e in range(self.ce):
            v = to_int(poly_eval(Matrix(shares[e].view(1, -1, 4), c), [0], np.stack([_mont(point)])))[0]
            acc.append(v * pow(y3, rank, P) % P)
        mine = torch.from_numpy(np.array([[(v >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)] for v in acc], dtype=np.uint64).view(np.int64)).to(dev)
        every = torch.empty((W,) + tuple(mine.shape), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(every, mine)
        per_rank = [to_int(a) for a in every.cpu().numpy().view(np.uint64)]
        total = [sum(vals) % P * rinv % P for vals in zip(*per_rank)]
        res.ood_trace, res.ood_composition = total[:len(taps)], total[len(taps):]
        self.mark("ood")
        # 12: DEEP quotient on the owned rows of the sub-coset, then extended like any other column
        coin.reseed_with_field_elements(res.ood_trace)
        coin.reseed_with_field_elements(res.ood_composition)
        alpha = res.deep_alpha = coin.draw()
        deep_prog = self.deep_template().patch([alpha], res.ood_trace + res.ood_composition, [0])
        reach = tap_reach(deep_prog.blob, log_N)
        for col, point in ((self.u_col, z), (self.v_col, zc)):
            lo_t, hi_t = reach.get(col, (0, 0))
            if hi_t - lo_t >= N // 4:
                inv_x_minus_c(all_lde[col], _mont(point), c, log_row_step=b)
                continue
            for lo, cnt in PN:
                inv_x_minus_c(all_lde[col], _mont(point), c, log_row_step=b, rows=((lo + lo_t) >> b, min(n, (cnt + hi_t - lo_t + (1 << b) - 1) >> b)))
        if self.value_col is not None:
            # the long pole sums as columns on the sub-coset rows, from sharded transform pairs (see _pole_sum_on_coset)
            v_w: dict[int, int] = {}
            col_w: dict[int, dict] = {col: {} for col in self.filter_cols}
            a_k = 1
            for (col, off), y in zip(taps, res.ood_trace):
                v_w[off] = (v_w.get(off, 0) + a_k * y) % P
                if col in col_w:
                    col_w[col][off] = (col_w[col].get(off, 0) + a_k) % P
                a_k = a_k * alpha % P
            self._pole_sums_on_coset([(v_w, all_lde[self.value_col, ::1 << b])] +
                                     [(col_w[col], all_lde[fcol, ::1 << b]) for col, fcol in self.filter_cols.items()], z)
        deep = torch.empty((N, 4), dtype=torch.int64, device=dev)
        quotient = torch.empty((n, 4), dtype=torch.int64, device=dev)
        self.mark("deep_setup")
        for lo, cnt in PN:
            evaluate(deep_prog, Matrix(all_lde, c), b, out=quotient, rows=(lo, cnt >> b), log_row_step=b)
        self.mark("deep")
        lde_one(quotient, deep, on_coset=True)
        self.mark("deep_lde")
        # 13: FRI.  Every rank needs the whole evaluation vector for the layers (their fold groups and leaf ranges do not follow
        #     the block-cyclic pieces): one all-gather, timed with the FRI stage it serves
        if capi:
            c.check(c.lib.ss_dist_allgather(c.handle, ctypes.c_void_p(deep.data_ptr()), log_N, _stream_ptr()))
        else:
            self._gather_pieces(deep, log_N)
        del comp_evals, quotient
        # 13: FRI layers on the gathered evaluations (tree-leaf / row ranges per rank while the layers are large)
        evals, log_size, offset = deep, log_N, 3
        layers = []
        while (1 << log_size) >> b > opt.max_remainder_coeffs and log_size > opt.log_fold:
            rows = 1 << (log_size - opt.log_fold)
            root, handle = self._commit(evals.data_ptr(), rows, 1 << opt.log_fold, log_size - opt.log_fold, _lib.ORDER_BITREV_RC)
            res.fri_roots.append(root)
            coin.reseed_with_digest(root)
            fri_alpha = coin.draw()
            res.fri_alphas.append(fri_alpha)
            shard = rows >= (1 << 16)
            lo, cnt = (rank * (rows // W), rows // W) if shard else (0, 0)
            nxt = fri_fold(evals, opt.log_fold, _mont(fri_alpha), _mont(offset), starkware_scale=True, ctx=c, rows=(lo, cnt) if shard else None)
            if shard:
                self._gather_rows(nxt, lo, cnt)
            layers.append((handle, self._last_subroots, evals, log_size))
            evals, log_size, offset = nxt, log_size - opt.log_fold, pow(offset, 1 << opt.log_fold, P)
        self.final_domain = (log_size, offset)
        res.remainder = self._remainder(evals, b)
        coin.reseed_with_field_element_vector([v * rinv % P for v in to_int(res.remainder)])
        self.mark("fri")
        if opt.grinding_factor:
            res.pow_nonce = coin.grind_proof_of_work(opt.grinding_factor, c)
            coin.reseed_with_int(res.pow_nonce)
        if queries:
            # 15: query phase.  Leaf p of a trace tree commits LDE row brev(p): the leaf's owner (contiguous leaf ranges) opens
            # the path, the row's owner (block-cyclic pieces) reads the values; both are shared, so every rank ends with the openings.
            pos = res.query_positions = coin.draw_queries(opt.num_queries, N)
            nat = [_brev(p, log_N) for p in pos]
            for name, (sub, subs, cols) in zip(("base", "ext", "composition"), trace_trees):
                paths = self._open_sharded(sub, subs, log_N, cols.shape[0], pos)
                rows_out = self._read_rows_sharded(cols, S, log_N, nat)
                res.opened_bytes += paths.nbytes + rows_out.nbytes
                if keep_openings:
                    res.trace_queries[name] = {"rows": rows_out, "paths": paths}
            F = 1 << opt.log_fold
            col_perm = [_brev(j, opt.log_fold) for j in range(F)]
            for handle, subs, layer_evals, ls in layers:
                log_rows = ls - opt.log_fold
                pos = sorted({p >> opt.log_fold for p in pos})
                paths = self._open_sharded(handle, subs, log_rows, F, pos)
                res.opened_bytes += paths.nbytes
                if keep_openings:                        # (every rank holds the whole layer: no exchange for the values)
                    idx = np.array([_brev(r, log_rows) for r in pos], dtype=np.uint64)
                    rows_out = np.zeros((len(idx), F, 4), dtype=np.uint64)
                    c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(layer_evals.data_ptr()), 1 << log_rows, F,
                                                 idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx), rows_out.ctypes.data_as(ctypes.c_void_p)))
                    res.fri_layers.append({"positions": pos, "rows": np.ascontiguousarray(rows_out[:, col_perm]), "paths": paths})
            self.mark("queries")
        for handle, _, _, _ in layers:
            c.lib.ss_tree_free(handle)
        for sub, _, _ in trace_trees:
            c.lib.ss_tree_free(sub)
        return res

    def stage_ms(self) -> dict:
        out = {}
        for (_, a), (name, b) in zip(self.timeline[:-1], self.timeline[1:]):
            if name != "start":
                out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out
