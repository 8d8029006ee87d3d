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
from .air.deep import deep_expr_shifted, deep_terms
from .air.evaluate import evaluate
from .air.expr import P
from .air.layouts import load_layout
from .matrix import Matrix, fri_fold, inv_x_minus_c, ood_eval, poly_eval

R = 2**256


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
    tree_kind: int = _lib.TREE_KECCAK_M20        # src/claims.rs:18-21 (starknet / EthVerifier); recursive claims use TREE_FRIENDLY
    n_friendly: int = 22                         # NUM_FRIENDLY_COMMITMENT_LAYERS, src/claims.rs:10 (TREE_FRIENDLY only)
    col_pad_rows: int = 0                        # padding between the columns of the working matrix (not a protocol parameter)


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
        self._template = None
        self._composition_program = None
        self._challenges = self._hints = self._alpha = None
        self.timeline: list = []

    # ---- what this rank reads of the trace (for callers that stream it from the host) -----------------------------
    def trace_columns_owned(self) -> list[int]:
        """trace columns (base then extension) whose LDE this rank computes: it needs them completely."""
        from .parallel import owned_columns

        nb, ne = self.layout.num_base_columns, self.layout.num_extension_columns
        return owned_columns(nb, self.rank, self.world) + [nb + j for j in owned_columns(ne, self.rank, self.world)]

    def trace_rows_needed(self) -> tuple[int, int]:
        """(first row, count) of the trace rows this rank reads of EVERY column (the out-of-domain dot products over its
        row range reach max_offset rows further; wraps mod n)."""
        step = self.n // self.world
        return self.rank * step, min(self.n, step + (self.layout.max_offset if self.world > 1 else 0))

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

    def composition_program(self, challenges, hints, alpha):
        """the per-proof value patch (a few milliseconds): constants <- challenges, hints, composition coefficient."""
        self._composition_program = self.composition_template().patch(challenges, hints, alpha)
        return self._composition_program

    # ---- commitment helper: whole tree on one GPU, row-range sub-trees + combined root on several ---------
    def _commit(self, ptr: int, col_stride: int, n_cols: int, log_rows: int):
        """Returns (root bytes, handle or None).  ptr: device address of column 0, row 0 of the matrix."""
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
        c.check(c.lib.ss_merkle_build(c.handle, opt.tree_kind, friendly, ctypes.c_void_p(ptr + 32 * lo), col_stride, n_cols,
                                      cnt.bit_length() - 1, _lib.ORDER_NATURAL, ctypes.byref(handle), None))
        root = (ctypes.c_uint8 * 32)()
        c.check(c.lib.ss_merkle_root(c.handle, handle, root))
        if shard:
            from .parallel import gather_subroots

            subs = gather_subroots(bytes(root), world, self.device)
            buf = (ctypes.c_uint8 * (32 * world)).from_buffer_copy(b"".join(subs))
            c.check(c.lib.ss_merkle_combine(c.handle, opt.tree_kind, buf, world.bit_length() - 1, root))
        return bytes(root), handle

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
        With world > 1 (torch.distributed initialised, one process per GPU): LDE and OOD are sharded by column,
        Merkle hashing / constraint evaluation / DEEP / FRI folds by LDE row range; LDE columns are broadcast
        from their owners, row-sharded vectors all-gathered, sub-tree roots combined (SURVEY.md §8e plan A)."""
        opt, L = self.opt, self.layout
        assert base.num_cols == L.num_base_columns and base.num_rows == self.n
        coin = self.coin
        from .parallel import owned_columns, share_row_ranges

        res = HotPathResult()
        dev, world, rank = self.device, self.world, self.rank
        n, N, b = self.n, self.N, opt.log_blowup
        c = self.ctx = base.ctx
        nb, C = L.num_base_columns, L.num_columns
        row_lo, row_cnt = rank * (N // world), N // world
        handles = []
        self.mark("start")
        # one matrix for every committed column: trace | composition (ce) | w | u | v  (see __init__)
        S = N + opt.col_pad_rows                         # column stride of the working matrix
        all_lde = torch.empty((C + self.ce + 3, S, 4), dtype=torch.int64, device=dev)[:, :N]
        lde = all_lde[:C]

        def lde_cols(src: Matrix, first_col: int):
            for j in owned_columns(src.num_cols, rank, world):
                if column_ready is not None:
                    column_ready(first_col + j)          # e.g. make the stream wait for the upload of this column
                c.check(c.lib.ss_lde(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src.data[j].data_ptr()), n, 1, self.log_n, b,
                                     ctypes.c_void_p(lde[first_col + j].data_ptr()), S, None, n,
                                     _lib.ORDER_NATURAL, None))

        # 3-5: base trace
        lde_cols(base, 0)
        self.mark("lde_base")
        halo = L.max_offset << b                      # forward reach of the constraint taps, in LDE rows
        share_row_ranges(lde[:nb], world, rank, halo)
        self.mark("share_base")
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
        share_row_ranges(lde[nb:], world, rank, halo)
        self.mark("share_ext")
        res.roots["ext"], h = self._commit(lde[nb].data_ptr(), S, C - nb, self.log_n + b); handles.append(h)
        self.mark("merkle_ext")
        coin.reseed_with_digest(res.roots["ext"])
        # 9: constraint evaluation (row range of this rank), boundary denominators from w = 1/(x - 1)
        res.composition_coeffs = [coin.draw()]
        prog = self.composition_program(challenges, res.hints, res.composition_coeffs)
        self.mark("patch")
        if world == 1:
            inv_x_minus_c(all_lde[self.w_col], _mont(1), c)
        else:                                            # only the rows this rank's taps reach
            lo_w, hi_w = tap_reach(prog.blob, self.log_n + b).get(self.w_col, (0, 0))
            if hi_w - lo_w >= N // 4:                    # tiny domains: signed offsets are ambiguous, take every row
                inv_x_minus_c(all_lde[self.w_col], _mont(1), c)
            else:
                inv_x_minus_c(all_lde[self.w_col], _mont(1), c, rows=(row_lo + lo_w, min(N, row_cnt + hi_w - lo_w)))
        comp_evals = torch.empty((N, 4), dtype=torch.int64, device=dev)
        self.mark("inv_w")
        evaluate(prog, Matrix(all_lde, c), b, out=comp_evals, rows=(row_lo, row_cnt) if world > 1 else None)
        self._gather_rows(comp_evals, row_lo, row_cnt)
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
        if world == 1:
            Matrix(comp_lde, c).ntt_(coset=True)
        else:                                            # one composition column per rank, then the row ranges
            for j in owned_columns(self.ce, rank, world):
                Matrix(comp_lde[j:j + 1], c).ntt_(coset=True)
            share_row_ranges(comp_lde, world, rank, 0)
        self.mark("ntt_comp_fwd")
        res.roots["composition"], h = self._commit(comp_lde.data_ptr(), S, self.ce, self.log_n + b); handles.append(h)
        self.mark("merkle_comp")
        # 11: out-of-domain evaluations of every tap, straight from the trace (barycentric dot products with one shared
        #     weight vector, ss_ood_eval).  Each rank sums over its range of trace rows; the partial values add up.
        coin.reseed_with_digest(res.roots["composition"])
        z = res.ood_point = coin.draw()
        taps = L.taps()
        if column_ready is not None:
            column_ready(None)                           # every trace column is read from here on
        t_lo, t_cnt = rank * (n // world), n // world
        parts = np.zeros((len(taps), 4), dtype=np.uint64)
        for mat, first, count in ((base, 0, nb), (ext, nb, C - nb)):
            idx = [k for k, (col, _) in enumerate(taps) if first <= col < first + count]
            if idx:
                parts[idx] = ood_eval(mat, [taps[k][0] - first for k in idx], [taps[k][1] for k in idx], _mont(z),
                                      rows=(t_lo, t_cnt) if world > 1 else None)
        to_int = lambda a: [int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192 for r in a]
        if world > 1:
            import torch.distributed as dist

            mine = torch.from_numpy(parts.view(np.int64)).to(dev)
            every = torch.empty((world,) + tuple(mine.shape), dtype=torch.int64, device=dev)
            dist.all_gather_into_tensor(every, mine)
            per_rank = [to_int(a) for a in every.cpu().numpy().view(np.uint64)]
            ood_m = [sum(vals) % P for vals in zip(*per_rank)]
        else:
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
        t_terms, c_terms = deep_terms(taps, res.ood_trace, res.ood_composition, self.comp_col, alpha, P)
        # (the sub-coset evaluation below reads u and v only at rows that are multiples of the blowup)
        if world == 1:                                   # (launched first: the GPU works while the host compiles)
            inv_x_minus_c(all_lde[self.u_col], _mont(z), c, log_row_step=b)
            inv_x_minus_c(all_lde[self.v_col], _mont(zc), c, log_row_step=b)
        deep_prog = compile_program(deep_expr_shifted(t_terms, c_terms, self.u_col, self.v_col, self.g, P), self.log_n, b)
        if world > 1:
            reach = tap_reach(deep_prog.blob, self.log_n + b)
            for col, point in ((self.u_col, z), (self.v_col, zc)):
                lo_t, hi_t = reach.get(col, (0, 0))
                first, cnt = (row_lo + lo_t) >> b, min(n, (row_cnt + hi_t - lo_t + (1 << b) - 1) >> b)
                if hi_t - lo_t >= N // 4:
                    first, cnt = 0, n
                inv_x_minus_c(all_lde[col], _mont(point), c, log_row_step=b, rows=(first, cnt))
        del comp_coeffs, comp_evals, work
        # The quotient has degree < n - 1, so its n values on the sub-coset 3<w_n> — the LDE rows that are multiples of
        # the blowup — determine it: evaluate only those (1/blowup of the work), then extend like any other column
        # (coset iNTT of size n, zero padding, coset NTT of size N).  Same polynomial, hence the same N evaluations.
        deep = torch.empty((N, 4), dtype=torch.int64, device=dev)
        sub_lo, sub_cnt = row_lo >> b, row_cnt >> b
        self.mark("deep_setup")
        evaluate(deep_prog, Matrix(all_lde, c), b, out=deep[:n], rows=(row_lo, sub_cnt) if world > 1 else None, log_row_step=b)
        self._gather_rows(deep[:n], sub_lo, sub_cnt)
        self.mark("deep")
        Matrix(deep[:n].view(1, n, 4), c).ntt_(inverse=True, coset=True)
        deep[n:].zero_()
        Matrix(deep.view(1, N, 4), c).ntt_(coset=True)
        self.mark("deep_lde")
        if self_check and world == 1:
            # test hook: the extended quotient equals the row-by-row evaluation on the whole LDE coset
            inv_x_minus_c(all_lde[self.u_col], _mont(z), c)
            inv_x_minus_c(all_lde[self.v_col], _mont(zc), c)
            res.deep_matches_full_evaluation = bool(torch.equal(deep, evaluate(deep_prog, Matrix(all_lde, c), b)))
        # 13: FRI layers
        evals, log_size, offset = deep, self.log_n + b, 3
        layers = []
        while (1 << log_size) >> b > opt.max_remainder_coeffs and log_size > opt.log_fold:
            rows = 1 << (log_size - opt.log_fold)
            # the layer matrix (rows x fold) is the evaluation buffer viewed with col_stride = rows
            root, handle = self._commit(evals.data_ptr(), rows, 1 << opt.log_fold, log_size - opt.log_fold)
            res.fri_roots.append(root)
            coin.reseed_with_digest(root)
            fri_alpha = coin.draw()
            res.fri_alphas.append(fri_alpha)
            shard = world > 1 and rows >= (1 << 16)
            lo, cnt = (rank * (rows // world), rows // world) if shard else (0, 0)
            nxt = fri_fold(evals, opt.log_fold, _mont(fri_alpha), _mont(offset), ctx=c, rows=(lo, cnt) if shard else None)
            if shard:
                self._gather_rows(nxt, lo, cnt)
            layers.append((handle, evals, log_size))
            evals, log_size, offset = nxt, log_size - opt.log_fold, pow(offset, 1 << opt.log_fold, P)
        res.remainder = evals.cpu().numpy().view(np.uint64)
        self.final_domain = (log_size, offset)
        coin.reseed_with_field_element_vector([v * rinv % P for v in to_int(res.remainder)])
        self.mark("fri")
        # 14: proof of work (GPU search for the smallest nonce), 15: query positions
        if opt.grinding_factor:
            res.pow_nonce = coin.grind_proof_of_work(opt.grinding_factor, c)
            coin.reseed_with_int(res.pow_nonce)
        # 15: queries (openings + rows through the ABI) — single-GPU trees only
        if queries and world == 1:
            pos = coin.draw_queries(opt.num_queries, N)
            res.query_positions = pos
            idx = np.array(pos, dtype=np.uint64)
            for name, h, (first, ncols) in zip(("base", "ext", "composition"), handles, ((0, nb), (nb, C - nb), (self.comp_col, self.ce))):
                paths = np.zeros((len(idx), self.log_n + b, 32), dtype=np.uint8)
                c.check(c.lib.ss_merkle_open(c.handle, h, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                             paths.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
                rows_out = np.zeros((len(idx), ncols, 4), dtype=np.uint64)
                c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(all_lde[first].data_ptr()), S, ncols,
                                             idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx), rows_out.ctypes.data_as(ctypes.c_void_p)))
                res.opened_bytes += paths.nbytes + rows_out.nbytes
                if keep_openings:
                    res.trace_queries[name] = {"rows": rows_out, "paths": paths}
            for handle, layer_evals, ls in layers:
                rows = 1 << (ls - opt.log_fold)
                idx = np.array(sorted({p % rows for p in pos}), dtype=np.uint64)
                paths = np.zeros((len(idx), ls - opt.log_fold, 32), dtype=np.uint8)
                c.check(c.lib.ss_merkle_open(c.handle, handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                             paths.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
                res.opened_bytes += paths.nbytes
                if keep_openings:
                    rows_out = np.zeros((len(idx), 1 << opt.log_fold, 4), dtype=np.uint64)
                    c.check(c.lib.ss_rows_gather(c.handle, ctypes.c_void_p(layer_evals.data_ptr()), rows, 1 << opt.log_fold,
                                                 idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx), rows_out.ctypes.data_as(ctypes.c_void_p)))
                    res.fri_layers.append({"positions": [int(p) for p in idx], "rows": rows_out, "paths": paths})
                pos = [int(p) for p in idx]
            self.mark("queries")
        for handle, _, _ in layers:
            c.lib.ss_tree_free(handle)
        for h in handles:
            c.lib.ss_tree_free(h)
        return res

    def stage_ms(self) -> dict:
        out = {}
        for (_, a), (name, b) in zip(self.timeline[:-1], self.timeline[1:]):
            if name != "start":
                out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out
