"""GPU hot path of `Stark::prove` — the stages of SURVEY.md §3.1 that run on the device, in the order
ministark's prover runs them (steps 3-5, 8-10, 11-13, 15), for a layout of the reference
(plain / recursive / starknet) and the proof options of the CLI (cli/src/main.rs:51-60):

    base trace  : LDE -> Merkle commit                      [challenges drawn]
    ext trace   : LDE -> Merkle commit                      [composition coefficient drawn]
    composition : constraint evaluation over the LDE coset -> coset iNTT -> ce interleaved columns
                  -> LDE -> Merkle commit                   [OOD point z drawn]
    OOD         : trace polynomials at z*g^offset for every tap, composition columns at z^ce
    DEEP        : sum alpha^i (T(x) - y) / (x - z g^k) over the LDE coset          (src/lib.rs:102-116)
    FRI         : per layer commit (rows of `fold` evaluations) -> alpha -> fold, until
                  layer_size / blowup <= max_remainder
    queries     : Merkle openings + rows at the query positions

Trace generation, `build_extension_columns`, the public coin / proof-of-work and proof serialisation
stay on the host in the reference (out of scope here), so challenges come from a seeded generator:
this class measures and checks the GPU stages, it does not emit a `Proof`."""
from __future__ import annotations

import ctypes
import random
from dataclasses import dataclass, field

import numpy as np
import torch

from . import _lib
from .air import compile_program
from .air.deep import deep_expr_shifted
from .air.evaluate import evaluate
from .air.expr import P
from .air.layouts import load_layout
from .matrix import Matrix, fri_fold, inv_x_minus_c, poly_eval
from .merkle import MatrixMerkleTree

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
    tree_kind: int = _lib.TREE_KECCAK_M20


@dataclass
class HotPathResult:
    roots: dict = field(default_factory=dict)
    fri_roots: list = field(default_factory=list)
    remainder: np.ndarray | None = None
    ood_trace: list = field(default_factory=list)
    ood_composition: list = field(default_factory=list)
    query_positions: list = field(default_factory=list)
    opened_bytes: int = 0


class HotPathProver:
    def __init__(self, layout: str, log_n: int, options: ProofOptions | None = None, seed: int = 0xB200, device=None):
        self.layout = load_layout(layout)
        self.log_n, self.opt = log_n, options or ProofOptions()
        self.n, self.N = 1 << log_n, 1 << (log_n + self.opt.log_blowup)
        self.ce = 1 << self.opt.log_blowup                      # ce_blowup_factor == lde blowup for Cairo (degree-2 constraints)
        self.rnd = random.Random(seed)
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self.g = pow(3, (P - 1) // self.n, P)
        # columns of the working matrix: trace | composition (ce) | w = 1/(x-1) | u = 1/(x-z) | v = 1/(x-z^ce)
        C = self.layout.num_columns
        self.comp_col, self.w_col, self.u_col, self.v_col = C, C + self.ce, C + self.ce + 1, C + self.ce + 2
        self._composition_program = None
        self._challenges = self._hints = self._alpha = None
        self.timeline: list = []

    # ---- host-side stand-ins for the public coin -------------------------------------------------------
    def _draw(self) -> int:
        return self.rnd.randrange(P)

    def mark(self, name: str):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        self.timeline.append((name, ev))

    def composition_program(self):
        """AirConfig::constraints + composition_constraint, compiled once per (layout, n, challenges)."""
        if self._composition_program is None:
            L = self.layout
            self._challenges = [self._draw() for _ in range(L.n_challenges())]
            self._hints = [self._draw() for _ in range(L.n_hints())]
            self._alpha = [self._draw()]
            # boundary denominators X - g^e read the auxiliary column w = 1/(x - 1) (see Layout.constraints)
            self._composition_program = compile_program(L.composition(self.n, inv_x_minus_one_col=self.w_col), self.log_n,
                                                        self.opt.log_blowup, self._challenges, self._hints, self._alpha)
        return self._composition_program

    # ---- the device stages ---------------------------------------------------------------------------------
    def prove(self, base: Matrix, ext: Matrix, queries: bool = True) -> HotPathResult:
        opt, L = self.opt, self.layout
        assert base.num_cols == L.num_base_columns and ext.num_cols == L.num_extension_columns and base.num_rows == self.n
        res = HotPathResult()
        dev = self.device
        n, N, b = self.n, self.N, opt.log_blowup
        self.mark("start")
        # 3-5: base trace
        # one matrix for every committed column: trace columns first, the ce composition columns last
        # (the DEEP stage reads all of them; a single allocation avoids a 50 GB concatenation at 2^22 steps)
        all_lde = torch.empty((L.num_columns + self.ce + 3, N, 4), dtype=torch.int64, device=dev)
        lde = all_lde[: L.num_columns]
        coeffs = torch.empty((L.num_columns, n, 4), dtype=torch.int64, device=dev)
        c = base.ctx
        nb = L.num_base_columns

        def lde_into(src: Matrix, first_col: int):
            c.check(c.lib.ss_lde(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src.data.data_ptr()), n, src.num_cols, self.log_n, b,
                                 ctypes.c_void_p(lde[first_col].data_ptr()), N, ctypes.c_void_p(coeffs[first_col].data_ptr()), n,
                                 _lib.ORDER_NATURAL, None))

        lde_into(base, 0)
        self.mark("lde_base")
        base_tree = MatrixMerkleTree.from_matrix(Matrix(lde[:nb], c), opt.tree_kind)
        self.mark("merkle_base")
        # 8: extension trace
        lde_into(ext, nb)
        self.mark("lde_ext")
        ext_tree = MatrixMerkleTree.from_matrix(Matrix(lde[nb:], c), opt.tree_kind)
        self.mark("merkle_ext")
        # 9: constraint evaluation
        prog = self.composition_program()
        inv_x_minus_c(all_lde[self.w_col], _mont(1), c)
        comp_evals = evaluate(prog, Matrix(all_lde, c), b)
        self.mark("constraint_eval")
        # 10: composition polynomial -> ce columns (coefficients j, j+ce, ...) -> LDE -> commit
        work = Matrix(comp_evals.view(1, N, 4), c)
        work.ntt_(inverse=True, coset=True)
        self.mark("ntt_comp_inv")
        comp_coeffs = comp_evals.view(n, self.ce, 4).permute(1, 0, 2).contiguous()        # [ce, n, 4] natural order
        comp_lde = all_lde[self.comp_col:self.comp_col + self.ce]
        comp_lde.zero_()
        comp_lde[:, :n] = comp_coeffs
        self.mark("comp_split")
        comp_lde_m = Matrix(comp_lde, c).ntt_(coset=True)
        self.mark("ntt_comp_fwd")
        comp_tree = MatrixMerkleTree.from_matrix(comp_lde_m, opt.tree_kind)
        self.mark("merkle_comp")
        res.roots = {"base": base_tree.root(), "ext": ext_tree.root(), "composition": comp_tree.root()}
        # 11: out-of-domain evaluations
        z = self._draw()
        taps = L.taps()
        pts = [z * pow(self.g, off, P) % P for _, off in taps]
        ood = poly_eval(Matrix(coeffs, c), [col for col, _ in taps], np.stack([_mont(p) for p in pts]))
        zc = pow(z, self.ce, P)
        ood_c = poly_eval(Matrix(comp_coeffs, c), list(range(self.ce)), np.stack([_mont(zc)] * self.ce), natural_order=True)
        self.mark("ood")
        from_m = lambda a: [(int(r[0]) | int(r[1]) << 64 | int(r[2]) << 128 | int(r[3]) << 192) * pow(R, -1, P) % P for r in a]
        res.ood_trace, res.ood_composition = from_m(ood), from_m(ood_c)
        # 12: DEEP composition over the LDE coset (coefficients = powers of one alpha, src/lib.rs:102-116)
        alpha = self._draw()
        t_terms, c_terms, k = [], [], 0
        for (col, off), y in zip(taps, res.ood_trace):
            t_terms.append((col, off, y, pow(alpha, k, P))); k += 1
        for j, y in enumerate(res.ood_composition):
            c_terms.append((self.comp_col + j, y, pow(alpha, k, P))); k += 1
        inv_x_minus_c(all_lde[self.u_col], _mont(z), c)
        inv_x_minus_c(all_lde[self.v_col], _mont(zc), c)
        deep_prog = compile_program(deep_expr_shifted(t_terms, c_terms, self.u_col, self.v_col, self.g, P), self.log_n, b)
        del coeffs, comp_coeffs, comp_evals, work
        deep = evaluate(deep_prog, Matrix(all_lde, c), b)
        self.mark("deep")
        # 13: FRI layers
        evals, log_size, offset = deep, self.log_n + b, 3
        layers = []
        while (1 << log_size) >> b > opt.max_remainder_coeffs and log_size > opt.log_fold:
            rows = 1 << (log_size - opt.log_fold)
            handle = ctypes.c_void_p()
            # the layer matrix (rows x fold) is the evaluation buffer viewed with col_stride = rows
            c.check(c.lib.ss_merkle_build(c.handle, opt.tree_kind, 0, ctypes.c_void_p(evals.data_ptr()), rows, 1 << opt.log_fold,
                                          log_size - opt.log_fold, _lib.ORDER_NATURAL, ctypes.byref(handle), None))
            root = (ctypes.c_uint8 * 32)()
            c.check(c.lib.ss_merkle_root(c.handle, handle, root))
            res.fri_roots.append(bytes(root))
            fri_alpha = self._draw()
            nxt = fri_fold(evals, opt.log_fold, _mont(fri_alpha), _mont(offset), ctx=c)
            layers.append((handle, evals, log_size))
            evals, log_size, offset = nxt, log_size - opt.log_fold, pow(offset, 1 << opt.log_fold, P)
        res.remainder = evals.cpu().numpy().view(np.uint64)
        self.final_domain = (log_size, offset)
        self.mark("fri")
        # 15: queries (positions from the seeded generator; openings + rows through the ABI)
        if queries:
            pos = sorted({self.rnd.randrange(N) for _ in range(opt.num_queries)})
            res.query_positions = pos
            for tree in (base_tree, ext_tree, comp_tree):
                pr = tree.prove_rows(pos)
                res.opened_bytes += pr["rows"].nbytes + pr["paths"].nbytes
            for handle, layer_evals, ls in layers:
                rows = 1 << (ls - opt.log_fold)
                idx = np.array(sorted({p % rows for p in pos}), dtype=np.uint64)
                paths = np.zeros((len(idx), ls - opt.log_fold, 32), dtype=np.uint8)
                c.check(c.lib.ss_merkle_open(c.handle, handle, idx.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(idx),
                                             paths.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8))))
                res.opened_bytes += paths.nbytes
                pos = [int(p) for p in idx]
            self.mark("queries")
        for handle, _, _ in layers:
            c.lib.ss_tree_free(handle)
        return res

    def stage_ms(self) -> dict:
        out = {}
        for (_, a), (name, b) in zip(self.timeline[:-1], self.timeline[1:]):
            if name != "start":
                out[name] = out.get(name, 0.0) + a.elapsed_time(b)
        return out
