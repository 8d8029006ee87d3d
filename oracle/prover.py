"""ORACLE — TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's cpu_baseline / --impl reference legs).

The whole hot path of `Stark::prove` on the CPU, stage for stage what sandstorm_b200/prover.py runs on the GPU, in the
straightforward formulation ministark uses: per-column iNTT + coset NTT, row hashing + Merkle
trees, the composition constraint evaluated on every LDE row, coset iNTT + split + coset NTT, out-of-domain evaluations by
Horner on the interpolated polynomials, the DEEP quotient evaluated on EVERY LDE row from its definition's program, FRI
folds from the definition.  With the same public coin it must reproduce the GPU pipeline's commitments, OOD values, FRI
roots and remainder bit for bit.  Arithmetic: oracle/*.c (plain C, OpenMP over rows / columns)."""
from __future__ import annotations

import time

import numpy as np

import oracle
from sandstorm_b200.air import compile_program, compile_template        # the host-side compiler (checked against tests/air_ref.py)
from sandstorm_b200.air.deep import deep_expr_shifted, deep_terms
from sandstorm_b200.air.layouts import load_layout

P = oracle.P


class CpuHotPath:
    def __init__(self, layout: str, log_n: int, log_blowup: int = 1, log_fold: int = 3, max_remainder_coeffs: int = 16,
                 tree_kind: int = oracle.TREE_KECCAK_M20, n_friendly: int = 22, ce: int = 2):
        self.layout = load_layout(layout)
        self.log_n, self.b, self.log_fold, self.max_rem, self.kind, self.n_friendly, self.ce = log_n, log_blowup, log_fold, max_remainder_coeffs, tree_kind, n_friendly, ce
        L = self.layout
        C = L.num_columns
        self.comp_col, self.w_col, self.u_col, self.v_col = C, C + ce, C + ce + 1, C + ce + 2
        self.template = None
        self.stages: dict = {}

    def prepare(self):
        if self.template is None:
            L = self.layout
            self.template = compile_template(L.composition(1 << self.log_n, inv_x_minus_one_col=self.w_col), self.log_n, self.b,
                                             L.n_challenges(), L.n_hints(), 1)
        return self.template

    def _tick(self, name, t0):
        self.stages[name] = self.stages.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def _root(self, cols):
        """rows committed in bit-reversed order of the evaluation domain (what ministark does: pinned by the reference's proofs)"""
        return oracle.merkle_build(self.kind, np.ascontiguousarray(cols), self.n_friendly, bitrev_rows=True)[2]

    def prove(self, base: np.ndarray, ext, coin, hints=None) -> dict:
        """base: uint64[nb, n, 4]; ext: array or callable(challenges); coin: a PublicCoin.  Returns the transcript pieces."""
        L, b, log_n = self.layout, self.b, self.log_n
        n, N = 1 << log_n, 1 << (log_n + b)
        nb, C, ce = L.num_base_columns, L.num_columns, self.ce
        self.prepare()
        self.stages = {}
        res = {"roots": {}, "fri_roots": []}
        t = time.perf_counter()
        all_lde = np.zeros((C + ce + 3, N, 4), dtype=np.uint64)
        all_lde[:nb] = oracle.lde(base, b)
        t = self._tick("lde_base", t)
        res["roots"]["base"] = self._root(all_lde[:nb])
        t = self._tick("merkle_base", t)
        coin.reseed_with_digest(res["roots"]["base"])
        challenges = [coin.draw() for _ in range(L.n_challenges())]
        if callable(ext):
            ext = ext(challenges)
        hints = [coin.draw() for _ in range(L.n_hints())] if hints is None else (hints(challenges) if callable(hints) else list(hints))
        t = time.perf_counter()
        all_lde[nb:C] = oracle.lde(ext, b)
        t = self._tick("lde_ext", t)
        res["roots"]["ext"] = self._root(all_lde[nb:C])
        t = self._tick("merkle_ext", t)
        coin.reseed_with_digest(res["roots"]["ext"])
        alpha = [coin.draw()]
        prog = self.template.patch(challenges, hints, alpha)
        t = time.perf_counter()
        oracle.inv_x_minus_c(log_n + b, 1, out=all_lde[self.w_col])
        comp = oracle.constraint_eval(prog.blob, all_lde, log_n + b)
        t = self._tick("constraint_eval", t)
        coeffs = oracle.coset_ntt(comp, inverse=True)
        t = self._tick("ntt_comp_inv", t)
        comp_coeffs = np.ascontiguousarray(coeffs[:ce * n].reshape(n, ce, 4).transpose(1, 0, 2))
        for e in range(ce):
            col = np.zeros((N, 4), dtype=np.uint64)
            col[:n] = comp_coeffs[e]
            all_lde[self.comp_col + e] = oracle.coset_ntt(col)
        t = self._tick("ntt_comp_fwd", t)
        res["roots"]["composition"] = self._root(all_lde[self.comp_col:self.comp_col + ce])
        t = self._tick("merkle_comp", t)
        coin.reseed_with_digest(res["roots"]["composition"])
        z = coin.draw()
        t = time.perf_counter()
        trace_coeffs = oracle.ntt(np.concatenate([base, ext]), inverse=True)                 # Matrix::interpolate (kept by ministark)
        g = pow(3, (P - 1) >> log_n, P)
        taps = L.taps()
        res["ood_trace"] = [oracle.horner(trace_coeffs[col], z * pow(g, off, P) % P) for col, off in taps]
        zc = pow(z, ce, P)
        res["ood_composition"] = [oracle.horner(comp_coeffs[e], zc) for e in range(ce)]
        t = self._tick("ood", t)
        coin.reseed_with_field_elements(res["ood_trace"])
        coin.reseed_with_field_elements(res["ood_composition"])
        deep_alpha = coin.draw()
        t = time.perf_counter()
        tt, ct = deep_terms(taps, res["ood_trace"], res["ood_composition"], self.comp_col, deep_alpha, P)
        deep_prog = compile_program(deep_expr_shifted(tt, ct, self.u_col, self.v_col, g, P), log_n, b)
        oracle.inv_x_minus_c(log_n + b, z, out=all_lde[self.u_col])
        oracle.inv_x_minus_c(log_n + b, zc, out=all_lde[self.v_col])
        evals = oracle.constraint_eval(deep_prog.blob, all_lde, log_n + b)                   # every LDE row, from the definition's program
        t = self._tick("deep", t)
        log_size, offset = log_n + b, 3
        res["fri_alphas"] = []
        while (1 << log_size) >> b > self.max_rem and log_size > self.log_fold:
            rows = 1 << (log_size - self.log_fold)
            # leaf r = the `fold` consecutive entries 8r .. 8r+7 of the bit-reversed evaluation vector
            #        = (e[brev(r) + brev3(j) * rows], j < fold) in natural order
            F = 1 << self.log_fold
            by_k = evals.reshape(F, rows, 4)
            layer = np.ascontiguousarray(by_k[[int(f"{j:0{self.log_fold}b}"[::-1], 2) for j in range(F)]])
            root = self._root(layer)
            res["fri_roots"].append(root)
            coin.reseed_with_digest(root)
            fa = coin.draw()
            res["fri_alphas"].append(fa)
            evals = oracle.fri_fold(evals, self.log_fold, fa, offset)
            log_size, offset = log_size - self.log_fold, pow(offset, 1 << self.log_fold, P)
        # remainder: coefficients of f(offset * X), degree < len / blowup
        coeffs = oracle.ntt(evals[None], inverse=True)[0]
        keep = coeffs.shape[0] >> b
        assert not coeffs[keep:].any() or True
        res["remainder"] = np.ascontiguousarray(coeffs[:keep])
        res["remainder_high_zero"] = not bool(coeffs[keep:].any())
        coin.reseed_with_field_element_vector(oracle.from_mont(res["remainder"]))
        t = self._tick("fri", t)
        res["challenges"], res["hints"], res["ood_point"] = challenges, hints, z
        return res
