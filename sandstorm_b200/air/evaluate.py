"""Runs a compiled composition-constraint program on a device-resident LDE matrix
(the GPU replacement of ministark's `AirConfig::eval_constraint`)."""
from __future__ import annotations

import ctypes

import torch

from .. import _lib
from ..matrix import Matrix, _stream_ptr
from .program import CompiledProgram


def evaluate(program: CompiledProgram, lde: Matrix, log_blowup: int, out: torch.Tensor | None = None,
             rows: tuple[int, int] | None = None, log_row_step: int = 0) -> torch.Tensor:
    """lde: all trace columns (base then extension) on the LDE coset, natural order.
    Returns int64[N >> log_row_step, 4]: the composition evaluations on the rows that are multiples of
    2^log_row_step.  rows = (begin, count) evaluates only `count` such rows starting at LDE row `begin` (one rank's
    share), written at position (row >> log_row_step) of `out`."""
    c = lde.ctx
    if out is None:
        out = torch.empty((lde.num_rows >> log_row_step, 4), dtype=torch.int64, device=lde.data.device)
    begin, count = rows if rows is not None else (0, 0)
    c.check(c.lib.ss_constraint_eval(c.handle, program.blob, len(program.blob), ctypes.c_void_p(lde.data.data_ptr()), lde.col_stride,
                                     lde.num_cols, lde.log_rows - log_blowup, log_blowup, begin, count, log_row_step,
                                     ctypes.c_void_p(out.data_ptr()), _stream_ptr()))
    return out
