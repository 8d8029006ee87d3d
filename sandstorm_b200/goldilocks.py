"""Goldilocks (p = 2^64 - 2^32 + 1) side of `Matrix::interpolate / evaluate`: the single-limb NTT path of
BASELINE config 4 (reference: cli/src/main.rs:103-124 wires `ministark_gpu::fields::p18446744069414584321::ark::Fp`).
Columns are int64 tensors [n_cols, n] holding the u64 words of ark-ff `Fp64<MontBackend<_, 1>>` (x * 2^64 mod p)."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from .context import Context, default_context
from .matrix import _stream_ptr

P = 2**64 - 2**32 + 1
GENERATOR = 7


def _check(data: torch.Tensor):
    if data.dtype != torch.int64 or data.dim() != 2 or not data.is_cuda or not data.is_contiguous():
        raise ValueError("expected a contiguous CUDA int64 tensor [n_cols, n]")
    n = data.shape[1]
    if n & (n - 1) or n == 0:
        raise ValueError("column length must be a power of two")
    return n.bit_length() - 1


def ntt_(data: torch.Tensor, inverse: bool = False, coset: bool = False, in_order: int = _lib.ORDER_NATURAL,
         out_order: int = _lib.ORDER_NATURAL, ctx: Context | None = None) -> torch.Tensor:
    """In place: ark-poly fft / ifft (coset offset 7) of every column (ss_ntt with SS_FIELD_GOLDILOCKS)."""
    log_n = _check(data)
    c = ctx or default_context(data.device.index)
    c.check(c.lib.ss_ntt(c.handle, _lib.FIELD_GOLDILOCKS, ctypes.c_void_p(data.data_ptr()), data.shape[1], data.shape[0], log_n,
                         int(inverse), int(coset), in_order, out_order, _stream_ptr()))
    return data


def lde(trace: torch.Tensor, log_blowup: int, out_order: int = _lib.ORDER_NATURAL, ctx: Context | None = None) -> torch.Tensor:
    """interpolate on <w_n> + evaluate on 7 * <w_N>, N = n << log_blowup (ss_lde with SS_FIELD_GOLDILOCKS)."""
    log_n = _check(trace)
    c = ctx or default_context(trace.device.index)
    n_cols, n = trace.shape
    out = torch.empty((n_cols, n << log_blowup), dtype=torch.int64, device=trace.device)
    c.check(c.lib.ss_lde(c.handle, _lib.FIELD_GOLDILOCKS, ctypes.c_void_p(trace.data_ptr()), n, n_cols, log_n, log_blowup,
                         ctypes.c_void_p(out.data_ptr()), n << log_blowup, None, n, out_order, _stream_ptr()))
    return out
