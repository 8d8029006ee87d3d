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


# =====================================================================================================================
# Row-sharded transforms (SURVEY.md §8e plan B): ONE column's NTT / LDE split over the W ranks.
#
# A transform of size M = W * m is factored into a size-W transform ACROSS the ranks and size-m transforms INSIDE a rank,
# with one all-to-all in between (s = m / W):
#
#   A (block-cyclic in -> cyclic out), x[j1 m + j2], rank r holds j2 in [r s, (r+1) s) of every block j1:
#       y[k1][j2] = w^(j2 k1) * sum_j1 x[j1 m + j2] w_W^(j1 k1)          local size-W DFT + twiddle   (ss_shard_dft)
#       all-to-all: rank k1 collects y[k1][0..m)
#       X[k1 + W k2] = sum_j2 y[k1][j2] (w^W)^(j2 k2)                     local size-m transform       (ss_ntt_shard, stage 1)
#   B (cyclic in -> block-cyclic out), rank j1 holds x[j1 + W j2]:
#       z[k2] = w^(j1 k2) * sum_j2 x[j1 + W j2] (w^W)^(j2 k2)             local size-m transform + twiddle (ss_ntt_shard, stage 2)
#       all-to-all: rank r collects z_j1[r s .. (r+1) s) of every j1
#       X[k1 m + k2] = sum_j1 z_j1[k2] w_W^(j1 k1)                        local size-W DFT             (ss_shard_dft)
#
# An LDE is A at size n with the inverse root followed by B at size N = b n; the coefficient scaling in between
# (1/n and the coset shift 3^j, j = j1 + W j2) is a geometric sequence in the local index, folded into the local
# transform.  The result is "block-cyclic": rank r owns, of every block k1 of N / W rows, the rows
# [k1 N/W + r N/W^2, k1 N/W + (r+1) N/W^2) — W contiguous pieces, with every column of a row on the same rank, which is
# what the row-wise consumers (leaf hashing, constraint evaluation, DEEP) need.  Two all-to-alls per LDE column, each
# moving (W-1)/W of 1/W of the column; every rank does exactly 1/W of the arithmetic whatever the number of columns.
# The local kernels are behind an `ops` object so that the same orchestration runs on CPU tensors with big-int
# arithmetic in the gloo test-suite (tests/test_parallel_gloo.py).
P252 = 2**251 + 17 * 2**192 + 1
GEN = 3


def pieces(log_len: int, rank: int, world: int) -> list[tuple[int, int]]:
    """(first row, count) of the W row ranges rank `rank` owns of a block-cyclic vector of 2^log_len rows."""
    m = (1 << log_len) // world
    s = m // world
    if s == 0:
        raise ValueError("vector too short for block-cyclic sharding (needs at least world^2 rows)")
    return [(k1 * m + rank * s, s) for k1 in range(world)]


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, world: int) -> None:
    """out[src] <- inp[dst] of rank src, for tensors of shape [world, ...]."""
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(out, inp)
        return
    rank = dist.get_rank()                       # gloo has no all-to-all: point-to-point sends
    out[rank] = inp[rank]
    ops = []
    for other in range(world):
        if other != rank:
            ops.append(dist.P2POp(dist.isend, inp[other], other))
            ops.append(dist.P2POp(dist.irecv, out[other], other))
    for req in dist.batch_isend_irecv(ops):
        req.wait()


class DeviceShardOps:
    """the local kernels through the C ABI (CUDA tensors)."""

    def __init__(self, ctx):
        self.ctx = ctx

    @staticmethod
    def _mont(v: int) -> bytes:
        return (v % P252 * (1 << 256) % P252).to_bytes(32, "little")

    def dft(self, src, src_off, src_stride, dst, dst_off, dst_stride, count, log_w, inverse, tw_log_m, tw_offset):
        import ctypes

        from . import _lib
        from .matrix import _stream_ptr

        c = self.ctx
        c.check(c.lib.ss_shard_dft(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src.data_ptr() + 32 * src_off), src_stride,
                                   ctypes.c_void_p(dst.data_ptr() + 32 * dst_off), dst_stride, count, log_w, int(inverse), tw_log_m,
                                   tw_offset, _stream_ptr()))

    def ntt_shard(self, src, log_m, stages, log_expand, c0, h0, tw, dst):
        import ctypes

        from . import _lib
        from .matrix import _stream_ptr

        c = self.ctx
        c.check(c.lib.ss_ntt_shard(c.handle, _lib.FIELD_FP252, ctypes.c_void_p(src.data_ptr()), 1 << log_m, 1, log_m, stages, log_expand,
                                   self._mont(c0) if c0 is not None else None, self._mont(h0) if h0 is not None else None,
                                   self._mont(tw) if tw is not None else None, ctypes.c_void_p(dst.data_ptr()),
                                   1 << (log_m + log_expand), _stream_ptr()))


class ShardedTransforms:
    """Transforms of single columns sharded over the ranks of the default process group.  Vectors are full-length
    tensors [len, 4] of which this rank owns the block-cyclic pieces (`pieces`); only owned rows are read or written."""

    def __init__(self, rank: int, world: int, ops, device):
        if world & (world - 1) or world > 8 or world < 2:
            raise ValueError("sharded transforms need 2, 4 or 8 ranks")
        self.rank, self.world, self.ops, self.device = rank, world, ops, device
        self.log_w = world.bit_length() - 1
        self._buf: dict = {}

    def _tmp(self, name: str, rows: int) -> torch.Tensor:
        t = self._buf.get(name)
        if t is None or t.shape[0] < rows:
            t = self._buf[name] = torch.empty((rows, 4), dtype=torch.int64, device=self.device)
        return t[:rows]

    def _twiddle(self, log_M: int):
        """w_M^rank: the factor z[k2] *= (w_M^rank)^k2 of transform B."""
        return pow(pow(GEN, (P252 - 1) >> log_M, P252), self.rank, P252) if self.rank else None

    # ---- A: block-cyclic evaluations -> this rank's residue class of the (scaled) coefficients, bit-reversed -------
    def to_coefficients(self, src: torch.Tensor, log_len: int, c0: int, h0: int, out: torch.Tensor | None = None) -> torch.Tensor:
        """src: evaluations on a size-2^log_len domain, block-cyclic.  Returns [m, 4]: position brev(k2) holds
        c0 * h0^k2 * C[rank + W k2], C the UNNORMALISED inverse transform (sum_j x[j] w^(-j k))."""
        W, r = self.world, self.rank
        m = (1 << log_len) // W
        s = m // W
        send, recv = self._tmp("send", m), self._tmp("recv", m)
        self.ops.dft(src, r * s, m, send, 0, s, s, self.log_w, True, log_len, r * s)
        _all_to_all(recv.view(W, s, 4), send.view(W, s, 4), W)
        out = out if out is not None else torch.empty((m, 4), dtype=torch.int64, device=self.device)
        self.ops.ntt_shard(recv, m.bit_length() - 1, 1, 0, c0, h0, None, out)
        return out

    # ---- B: this rank's residue class of the coefficients (bit-reversed) -> block-cyclic evaluations ------------------
    def from_coefficients(self, coeffs: torch.Tensor, log_m: int, log_expand: int, dst: torch.Tensor) -> None:
        """coeffs: [2^log_m, 4], position brev(j2) holds x[rank + W j2] (already coset-scaled).  dst: full-length vector of
        W * 2^(log_m + log_expand) rows; its owned pieces receive sum_j x[j] w_M^(j k)."""
        W, r = self.world, self.rank
        mN = 1 << (log_m + log_expand)
        sN = mN // W
        z = self._tmp("z", mN)
        self.ops.ntt_shard(coeffs, log_m, 2, log_expand, None, None, self._twiddle(log_m + log_expand + self.log_w), z)
        recv = self._tmp("recv2", mN)
        _all_to_all(recv.view(W, sN, 4), z.view(W, sN, 4), W)
        self.ops.dft(recv, 0, sN, dst, r * sN, mN, sN, self.log_w, False, -1, 0)

    # ---- LDE of one column: block-cyclic evaluations on <w_n> (or 3<w_n>) -> block-cyclic evaluations on 3<w_N> ---------
    def lde(self, src: torch.Tensor, log_n: int, log_blowup: int, dst: torch.Tensor, src_on_coset: bool = False) -> None:
        """src on <w_n> (trace columns) or, with src_on_coset, on 3<w_n> (the DEEP quotient); dst on 3<w_N>."""
        self.lde_begin(src, log_n)
        self.lde_finish(log_n, log_blowup, dst, src_on_coset)

    def lde_begin(self, src: torch.Tensor, log_n: int, slot: int = 0) -> None:
        """first half: the size-W transform of the owned rows and the first all-to-all (into this slot's buffers)."""
        W, r = self.world, self.rank
        m = (1 << log_n) // W
        s = m // W
        send, recv = self._tmp(f"send{slot}", m), self._tmp(f"recv{slot}", m)
        self.ops.dft(src, r * s, m, send, 0, s, s, self.log_w, True, log_n, r * s)
        _all_to_all(recv.view(W, s, 4), send.view(W, s, 4), W)

    def lde_finish(self, log_n: int, log_blowup: int, dst: torch.Tensor, src_on_coset: bool = False, slot: int = 0, scale=None) -> None:
        """second half: the local LDE, the second all-to-all and the size-W transform into the owned pieces of dst.
        scale = (c, h): coefficient k of the UNNORMALISED inverse transform is multiplied by c * h^k instead of the LDE's
        3^k / n (a transform pair onto another coset, e.g. the pole sums of the DEEP quotient)."""
        W, r = self.world, self.rank
        n = 1 << log_n
        m = n // W
        ninv = pow(n, -1, P252)
        c0, h0 = (ninv, 1) if src_on_coset else (ninv * pow(GEN, r, P252) % P252, pow(GEN, W, P252))
        if scale is not None:
            c0, h0 = scale[0] * pow(scale[1], r, P252) % P252, pow(scale[1], W, P252)
        recv = self._tmp(f"recv{slot}", m)
        mN = m << log_blowup
        sN = mN // W
        z = self._tmp(f"z{slot}", mN)
        self.ops.ntt_shard(recv, m.bit_length() - 1, 3, log_blowup, c0, h0, self._twiddle(log_n + log_blowup), z)
        recv2 = self._tmp(f"recv2{slot}", mN)
        _all_to_all(recv2.view(W, sN, 4), z.view(W, sN, 4), W)
        self.ops.dft(recv2, 0, sN, dst, r * sN, mN, sN, self.log_w, False, -1, 0)

    # ---- composition polynomial: coset evaluations (size N) -> ce = 2 interleaved coefficient columns -> their LDEs ------
    def composition_columns(self, evals: torch.Tensor, log_n: int, log_blowup: int, dst_cols) -> list:
        """evals: composition evaluations on 3<w_N>, block-cyclic.  The coefficient of X^j goes to column j mod 2 as its
        coefficient j // 2 (ministark: Matrix::from_rows(coeffs.chunks(ce))); each column (degree < n) is evaluated on
        3<w_N> into dst_cols[e] (block-cyclic).  Returns [column 0 share, column 1 share]: this rank's residue class
        i = rank + W j2 of each column's coefficients, COSET-SCALED (h[i] * 3^i) at position brev(j2) — the format
        ss_poly_eval reads, for the out-of-domain values of the composition columns."""
        W, r = self.world, self.rank
        log_N = log_n + log_blowup
        Ninv = pow(1 << log_N, -1, P252)
        g_inv = pow(GEN, -1, P252)
        # C[j], j = r + W k2, becomes coefficient i = (j - e) / 2 of column e = j mod 2 = r mod 2 (W is even), to be
        # scaled by 3^i for the coset evaluation: 3^-j / N * 3^i = 3^-ceil(r/2) / N * (3^-(W/2))^k2
        coeffs = self.to_coefficients(evals, log_N, Ninv * pow(g_inv, (r + 1) // 2, P252) % P252, pow(g_inv, W // 2, P252))
        # degree < 2n: of the N / W coefficients k2 only k2 < 2n / W can be non-zero — in bit-reversed order the positions
        # that are multiples of N / 2n (what the single-GPU path keeps as well)
        keep = coeffs[:: 1 << (log_blowup - 1)] if log_blowup > 1 else coeffs
        half = keep.shape[0] // 2
        # i = (r >> 1) + (W/2) k2 lives on rank i mod W at local index k2 >> 1: even k2 (the first half of a bit-reversed
        # array) go to rank r >> 1, odd k2 (second half) to rank (r >> 1) + W/2; both as column r mod 2.
        parts = [keep[:half].contiguous(), keep[half:].contiguous()]
        dests = (r >> 1, (r >> 1) + W // 2)
        got = [torch.empty((half, 4), dtype=torch.int64, device=self.device) for _ in range(2)]
        srcs = [2 * (r % (W // 2)) + e for e in range(2)]           # who holds my share of column e
        ops = []
        for par, d in enumerate(dests):
            if d == r:
                got[r & 1].copy_(parts[par])
            else:
                ops.append(dist.P2POp(dist.isend, parts[par], d))
        for e, src in enumerate(srcs):
            if src != r:
                ops.append(dist.P2POp(dist.irecv, got[e], src))
        for req in (dist.batch_isend_irecv(ops) if ops else []):
            req.wait()
        log_m = half.bit_length() - 1                               # log2(n / W)
        for e in range(2):
            self.from_coefficients(got[e], log_m, log_blowup, dst_cols[e])
        return got
