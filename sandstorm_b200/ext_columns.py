"""`Trace::build_extension_columns` on the device.

The reference computes the permutation / aggregation columns of the Cairo layouts on one CPU thread
(layouts/src/recursive/trace.rs:699-814 "TODO: multithread", starknet/trace.rs:997-1100, plain/trace.rs:277-329):
running products of `z - (alpha * value + address)` terms over the program-order and address-order memory
columns, of `z - value` terms over the unordered / ordered range-check and diluted-check cells, batch-inverted
and multiplied, plus the diluted-check aggregation recurrence.  Here each of them is one device prefix scan
(`ss_perm_product`, `ss_diluted_aggregate`: csrc/ext_columns.cu) over the base trace that is already resident
on the GPU, so the host never sees the 2^26-row columns between the two commitments.

The per-layout tables below restate where each virtual column lives: (column, row shift, row stride) from the
column enums of layouts/src/{plain,recursive,starknet}/air.rs."""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib
from .matrix import Matrix, _stream_ptr

P = 2**251 + 17 * 2**192 + 1
R = 2**256

# perms: (numerator (col, addr shift, value shift | None), denominator (...), row stride, z challenge, alpha challenge | None,
#         output (extension column, row shift))          agg: ((col, shift), stride, z, alpha, output (ext column, shift))
LAYOUTS = {
    # recursive/air.rs: Npc col 3 (:1499), Mem col 4 (:1589), RangeCheck col 5 OffDst 0 / Ordered 2 (:1637), DilutedCheck cols 1 / 2
    # (:1613), Permutation Memory (9, 0) RangeCheck (9, 1) DilutedCheck (8, 0) (:1697), aggregate col 7
    "recursive": dict(perms=[((3, 0, 1), (4, 0, 1), 2, 0, 1, (2, 0)), ((5, 0, None), (5, 2, None), 4, 2, None, (2, 1)),
                             ((1, 0, None), (2, 0, None), 1, 3, None, (1, 0))],
                      agg=((2, 0), 1, 4, 5, (0, 0))),
    # starknet/air.rs: Npc col 5, Mem col 6, RangeCheck col 7 (:3159), DilutedCheck col 7 Unordered 1 / Ordered 5, step 8 (:3137),
    # Permutation col 9: Memory 0, RangeCheck 1, DilutedCheck 7 (:3220), Aggregate 3
    "starknet": dict(perms=[((5, 0, 1), (6, 0, 1), 2, 0, 1, (0, 0)), ((7, 0, None), (7, 2, None), 4, 2, None, (0, 1)),
                            ((7, 1, None), (7, 5, None), 8, 3, None, (0, 7))],
                     agg=((7, 5), 8, 4, 5, (0, 3))),
    # plain/air.rs: Npc col 1, Mem col 2, RangeCheck col 3 (:719), Permutation col 5: Memory 0, RangeCheck 1 (:772)
    "plain": dict(perms=[((1, 0, 1), (2, 0, 1), 2, 0, 1, (0, 0)), ((3, 0, None), (3, 2, None), 4, 2, None, (0, 1))], agg=None),
}
NUM_EXT = {"recursive": 3, "starknet": 1, "plain": 1}


def _mont_bytes(v: int) -> bytes:
    return (v % P * R % P).to_bytes(32, "little")


def build_extension_columns(layout: str, base: Matrix, challenges) -> Matrix:
    """base: the base trace on the device (column-major, natural order); challenges: canonical ints in the order of the
    layout's challenge enums (Memory z, alpha; RangeCheck z; DilutedCheck perm z; aggregation z, alpha).  Returns the
    extension columns as a device Matrix, zero outside the cells the layout defines (as the reference leaves them)."""
    spec = LAYOUTS[layout]
    c, n = base.ctx, base.num_rows
    ext = torch.zeros((NUM_EXT[layout], n, 4), dtype=torch.int64, device=base.data.device)
    stride_b = base.col_stride

    def cell(col, shift):
        return ctypes.c_void_p(base.data.data_ptr() + 32 * (col * stride_b + shift)) if shift is not None else None

    for (ncol, na, nv), (dcol, da, dv), step, zi, ai, (ocol, oshift) in spec["perms"]:
        c.check(c.lib.ss_perm_product(c.handle, _lib.FIELD_FP252, cell(ncol, na), cell(ncol, nv), cell(dcol, da), cell(dcol, dv), step, n // step,
                                      _mont_bytes(challenges[zi]), _mont_bytes(challenges[ai]) if ai is not None else None,
                                      ctypes.c_void_p(ext[ocol].data_ptr() + 32 * oshift), step, _stream_ptr()))
    if spec["agg"] is not None:
        (col, shift), step, zi, ai, (ocol, oshift) = spec["agg"]
        c.check(c.lib.ss_diluted_aggregate(c.handle, _lib.FIELD_FP252, cell(col, shift), step, n // step, _mont_bytes(challenges[zi]),
                                           _mont_bytes(challenges[ai]), ctypes.c_void_p(ext[ocol].data_ptr() + 32 * oshift), step, _stream_ptr()))
    return Matrix(ext, c)
