"""Device-resident column-major matrix: the mirror of ``ministark::Matrix<Fp>``
(constructed at reference layouts/src/recursive/trace.rs:652-660; read through
num_rows / num_cols / read_row at crypto/src/merkle/utils.rs:20,37,40)."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .context import Context, default_context


def _stream_ptr() -> ctypes.c_void_p:
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


class Matrix:
    """``data``: torch.int64 CUDA tensor of shape (n_cols, n_rows, 4): Montgomery limbs, column-major."""

    def __init__(self, data: torch.Tensor, ctx: Context | None = None):
        if data.dtype != torch.int64 or data.dim() != 3 or data.shape[2] != 4 or not data.is_cuda:
            raise ValueError("Matrix expects a CUDA int64 tensor of shape (n_cols, n_rows, 4)")
        # columns may be padded: any column stride >= the number of rows is fine as long as each column is contiguous
        if not (data.stride(2) == 1 and data.stride(1) == 4 and data.stride(0) % 4 == 0 and data.stride(0) >= 4 * data.shape[1]):
            data = data.contiguous()
        n = data.shape[1]
        if n & (n - 1):
            raise ValueError("number of rows must be a power of two")
        self.data = data
        self.ctx = ctx or default_context(data.device.index)

    # ---- construction / export -------------------------------------------------------------
    @classmethod
    def from_numpy(cls, cols: np.ndarray, device: str = "cuda", ctx: Context | None = None) -> "Matrix":
        a = np.ascontiguousarray(cols, dtype=np.uint64)
        return cls(torch.from_numpy(a.view(np.int64)).to(device), ctx)

    def numpy(self) -> np.ndarray:
        return self.data.cpu().numpy().view(np.uint64)

    @property
    def num_cols(self) -> int:
        return self.data.shape[0]

    @property
    def num_rows(self) -> int:
        return self.data.shape[1]

    @property
    def col_stride(self) -> int:
        """elements between consecutive columns (>= num_rows; the `col_stride` argument of the C ABI)."""
        return self.data.stride(0) // 4 if self.data.shape[0] > 1 else max(self.num_rows, self.data.stride(0) // 4)

    @property
    def log_rows(self) -> int:
        return self.num_rows.bit_length() - 1

    def clone(self) -> "Matrix":
        return Matrix(self.data.clone(), self.ctx)

    # ---- ministark Matrix::{interpolate, evaluate} --------------------------------------------
    def ntt_(self, inverse: bool = False, coset: bool = False, in_order: int = _lib.ORDER_NATURAL,
             out_order: int = _lib.ORDER_NATURAL) -> "Matrix":
        c = self.ctx
        c.check(c.lib.ss_ntt(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(self.data.data_ptr()), self.col_stride,
                             self.num_cols, self.log_rows, int(inverse), int(coset), in_order, out_order, _stream_ptr()))
        return self

    def interpolate(self) -> "Matrix":
        """Matrix::interpolate(trace_domain): evaluations on <w_n> -> coefficients (natural order)."""
        return self.clone().ntt_(inverse=True)

    def evaluate(self, log_blowup: int = 0, coset: bool = True) -> "Matrix":
        """Matrix::evaluate(domain): coefficients -> evaluations on (3 if coset)*<w_N>, N = n << log_blowup."""
        n_cols, n = self.num_cols, self.num_rows
        out = torch.zeros((n_cols, n << log_blowup, 4), dtype=torch.int64, device=self.data.device)
        out[:, :n] = self.data
        return Matrix(out, self.ctx).ntt_(inverse=False, coset=coset)

    def lde(self, log_blowup: int, keep_coeffs: bool = False, out_order: int = _lib.ORDER_NATURAL):
        """interpolate + evaluate on the LDE coset, fused (ss_lde).  Returns lde or (lde, coeffs)."""
        c = self.ctx
        n_cols, n = self.num_cols, self.num_rows
        out = torch.empty((n_cols, n << log_blowup, 4), dtype=torch.int64, device=self.data.device)
        coeffs = torch.empty_like(self.data) if keep_coeffs else None
        c.check(c.lib.ss_lde(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(self.data.data_ptr()), self.col_stride, n_cols, self.log_rows,
                             log_blowup, ctypes.c_void_p(out.data_ptr()), n << log_blowup,
                             ctypes.c_void_p(coeffs.data_ptr()) if keep_coeffs else None, n, out_order, _stream_ptr()))
        lde = Matrix(out, c)
        return (lde, Matrix(coeffs, c)) if keep_coeffs else lde


# ---- FRI / OOD helpers (ministark FriProver::build_layers, DeepPolyComposer inputs) --------------
def _felt_bytes(value_mont_limbs) -> bytes:
    return np.ascontiguousarray(value_mont_limbs, dtype=np.uint64).tobytes()


def fri_fold(evals: torch.Tensor, log_fold: int, alpha_mont: np.ndarray, offset_mont: np.ndarray, starkware_scale: bool = False,
             ctx: Context | None = None, out: torch.Tensor | None = None, rows: tuple[int, int] | None = None) -> torch.Tensor:
    """evals: int64[N, 4] on the coset offset*<w_N> (natural order) -> int64[N >> log_fold, 4].
    rows = (begin, count): fold only those outputs (the row range of one rank), written at their absolute position."""
    ctx = ctx or default_context(evals.device.index)
    n = evals.shape[0]
    if out is None:
        out = torch.empty((n >> log_fold, 4), dtype=torch.int64, device=evals.device)
    a, h = _felt_bytes(alpha_mont), _felt_bytes(offset_mont)
    begin, count = rows if rows is not None else (0, 0)
    ctx.check(ctx.lib.ss_fri_fold(ctx.handle, _lib.FIELD_FP252, ctypes.c_void_p(evals.data_ptr()), n.bit_length() - 1, log_fold,
                                  a, h, int(starkware_scale), begin, count, ctypes.c_void_p(out.data_ptr()), _stream_ptr()))
    return out


def poly_eval(coeffs: "Matrix", cols, points_mont: np.ndarray, natural_order: bool = False) -> np.ndarray:
    """coeffs: the coefficient matrix returned by Matrix.lde(keep_coeffs=True) (or, with natural_order, plain
    natural-order coefficients); evaluates column cols[e] at points_mont[e] (Montgomery limbs).
    Returns uint64[len, 4]."""
    ctx = coeffs.ctx
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    pts = np.ascontiguousarray(points_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros_like(pts)
    torch.cuda.current_stream().synchronize()
    ctx.check(ctx.lib.ss_poly_eval(ctx.handle, _lib.FIELD_FP252, ctypes.c_void_p(coeffs.data.data_ptr()), coeffs.col_stride, coeffs.log_rows,
                                   int(natural_order), cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), pts.ctypes.data_as(ctypes.c_void_p), len(cols),
                                   out.ctypes.data_as(ctypes.c_void_p)))
    return out


def ood_eval(trace: "Matrix", cols, offsets, z_mont: np.ndarray, rows: tuple[int, int] | None = None) -> np.ndarray:
    """Out-of-domain values from the trace evaluations (ss_ood_eval, barycentric form): returns uint64[len, 4] with
    T_{cols[e]}(z * g^offsets[e]) in Montgomery limbs.  rows = (begin, count) restricts the sum to that range of trace
    rows: the results of disjoint ranges add up (mod p) to the value."""
    ctx = trace.ctx
    cols = np.ascontiguousarray(cols, dtype=np.int32)
    offs = np.ascontiguousarray([int(o) % trace.num_rows for o in offsets], dtype=np.uint64)
    out = np.zeros((len(cols), 4), dtype=np.uint64)
    begin, count = rows if rows is not None else (0, 0)
    torch.cuda.current_stream().synchronize()
    ctx.check(ctx.lib.ss_ood_eval(ctx.handle, _lib.FIELD_FP252, ctypes.c_void_p(trace.data.data_ptr()), trace.col_stride, trace.log_rows,
                                  cols.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), offs.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)), len(cols),
                                  _felt_bytes(z_mont), begin, count, out.ctypes.data_as(ctypes.c_void_p)))
    return out


def inv_x_minus_c(out: torch.Tensor, c_mont: np.ndarray, ctx: Context | None = None, log_row_step: int = 0,
                  rows: tuple[int, int] | None = None) -> torch.Tensor:
    """out[i] = 1 / (3 * w_N^i - c) for the N = out.shape[0] points of the LDE coset (ss_inv_x_minus_c); with
    log_row_step only the rows that are multiples of 2^log_row_step are computed and written; rows = (begin, count)
    in units of 2^log_row_step rows (begin may be negative: the range wraps mod N) restricts the work to one
    rank's row range plus the halo its shifted reads need."""
    ctx = ctx or default_context(out.device.index)
    n = out.shape[0]
    begin, count = rows if rows is not None else (0, 0)
    ctx.check(ctx.lib.ss_inv_x_minus_c(ctx.handle, _lib.FIELD_FP252, n.bit_length() - 1, log_row_step, begin % (n >> log_row_step), count, _felt_bytes(c_mont),
                                       ctypes.c_void_p(out.data_ptr()), _stream_ptr()))
    return out
