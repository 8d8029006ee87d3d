"""Multi-GPU plumbing for the hot path (plan A of BASELINE north_star): one process per
GPU, torch.distributed for the exchange.  The functions are backend-agnostic (NCCL on GPUs, gloo in the
CPU test-suite): they only decide who owns what and move whole columns / 32-byte sub-roots.

    LDE      : column j is transformed by rank j % world            (no communication)
    exchange : every LDE column is broadcast from its owner          (NCCL over NVLink)
    Merkle   : rank r hashes rows [r*N/world, (r+1)*N/world) and builds that sub-tree;
               the world sub-roots are all-gathered (32 B each) and combined (ss_merkle_combine)
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def owned_columns(n_cols: int, rank: int, world: int) -> list[int]:
    return [j for j in range(n_cols) if j % world == rank]


def owner_of(col: int, world: int) -> int:
    return col % world


def row_range(n_rows: int, rank: int, world: int) -> tuple[int, int]:
    if world & (world - 1) or n_rows % world:
        raise ValueError("world size must be a power of two dividing the row count")
    step = n_rows // world
    return rank * step, (rank + 1) * step


def share_columns(matrix: torch.Tensor, world: int) -> None:
    """In place: after the call every rank holds every column.  matrix: [n_cols, rows, limbs]."""
    if world == 1:
        return
    for j in range(matrix.shape[0]):
        dist.broadcast(matrix[j], src=owner_of(j, world))


def gather_subroots(my_root: bytes, world: int, device) -> list[bytes]:
    """All-gather of the per-rank sub-tree roots, in rank (= row) order."""
    if world == 1:
        return [my_root]
    mine = torch.tensor(list(my_root), dtype=torch.uint8, device=device)
    out = [torch.empty(32, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(r.cpu().numpy()) for r in out]
