"""Multi-GPU plumbing for the hot path (plan A of BASELINE north_star): one process per
GPU, torch.distributed for the exchange.  The functions are backend-agnostic (NCCL on GPUs, gloo in the
CPU test-suite): they only decide who owns what and move whole columns / 32-byte sub-roots.

    LDE      : column j is transformed by rank j % world            (no communication)
    exchange : every rank receives, from the owner of each column, the rows it will consume: its own row range
               plus a halo of max_offset * blowup rows for the constraint taps (share_row_ranges; point-to-point
               sends grouped into one NCCL launch — 1/world of the traffic of broadcasting whole columns)
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


def share_row_ranges(matrix: torch.Tensor, world: int, rank: int, halo: int) -> None:
    """In place: after the call rank r holds, for EVERY column, the rows [r*step, (r+1)*step + halo) (mod N) — what the
    row-sharded consumers read (Merkle leaves, constraint evaluation with its forward taps, DEEP).  Column j is
    complete on its owner before the call.  matrix: [n_cols, N, limbs]."""
    if world == 1:
        return
    n_cols, N = matrix.shape[0], matrix.shape[1]
    step = N // world
    if halo > step:                                   # tiny domains: the halo would span several ranks
        share_columns(matrix, world)
        return
    ops = []
    for j in range(n_cols):
        o = owner_of(j, world)
        for r in range(world):
            if r == o:
                continue
            rows = matrix[j, r * step:(r + 1) * step]
            if rank == o:
                ops.append(dist.P2POp(dist.isend, rows, r))
            elif rank == r:
                ops.append(dist.P2POp(dist.irecv, rows, o))
    for req in (dist.batch_isend_irecv(ops) if ops else []):
        req.wait()
    if halo == 0:
        return
    # halo: the `halo` rows after my range are the first rows of the next rank's range (wrapping at N)
    nxt, prv = (rank + 1) % world, (rank - 1) % world
    lo = ((rank + 1) * step) % N
    ops = []
    for j in range(n_cols):
        ops.append(dist.P2POp(dist.isend, matrix[j, rank * step:rank * step + halo], prv))
        ops.append(dist.P2POp(dist.irecv, matrix[j, lo:lo + halo], nxt))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


def redistribute(column: torch.Tensor, world: int, rank: int, chunk: int, to_cyclic: bool) -> None:
    """Ownership exchange of the row-sharded four-step NTT (tools/ntt_model.py), in place on a full-size
    column [N, limbs] of which this rank owns
        contiguous   : positions p with p // (N / world) == rank, or
        chunk-cyclic : positions p with (p // chunk) % world == rank.
    to_cyclic=True turns the first into the second (before the strided passes), False the reverse (before / after the
    contiguous pass).  Viewed as [world, K, world, chunk, limbs] the two maps are the first and the third axis, so the
    exchange is a block transpose: rank r sends x[r, :, d] to d and receives x[s, :, r] from s (and the mirror image)."""
    if world == 1:
        return
    N = column.shape[0]
    if N % (world * world * chunk):
        raise ValueError("column length must be a multiple of world^2 * chunk")
    x = column.view(world, N // (world * world * chunk), world, chunk, *column.shape[1:])
    sends, recvs, ops = [], [], []
    for other in range(world):
        if other == rank:
            continue
        src = x[rank, :, other] if to_cyclic else x[other, :, rank]
        buf = src.contiguous()
        landing = torch.empty_like(buf)
        sends.append(buf)
        recvs.append((other, landing))
        ops.append(dist.P2POp(dist.isend, buf, other))
        ops.append(dist.P2POp(dist.irecv, landing, other))
    for req in dist.batch_isend_irecv(ops):
        req.wait()
    for other, landing in recvs:
        if to_cyclic:
            x[other, :, rank] = landing
        else:
            x[rank, :, other] = landing


def gather_subroots(my_root: bytes, world: int, device) -> list[bytes]:
    """All-gather of the per-rank sub-tree roots, in rank (= row) order."""
    if world == 1:
        return [my_root]
    mine = torch.tensor(list(my_root), dtype=torch.uint8, device=device)
    out = [torch.empty(32, dtype=torch.uint8, device=device) for _ in range(world)]
    dist.all_gather(out, mine)
    return [bytes(r.cpu().numpy()) for r in out]
