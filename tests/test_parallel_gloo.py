"""CPU-only, world_size 2 over gloo: the sharding plan used at N > 1 GPUs (sandstorm_b200/parallel.py)
reproduces the single-process commitment: column-sharded LDE + broadcast == full LDE, and the combined
row-range sub-roots == the root of the whole tree.  The heavy arithmetic is done by the oracle here
(the product kernels need a GPU); what is under test is ownership, exchange and ordering."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, cols_np, q):
    import oracle
    from sandstorm_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n_cols, n, _ = cols_np.shape
        N = 2 * n
        lde = torch.zeros((n_cols, N, 4), dtype=torch.int64)
        for j in parallel.owned_columns(n_cols, rank, world):
            lde[j] = torch.from_numpy(oracle.lde(cols_np[j:j + 1], 1)[0].view(np.int64))
        # row-range exchange (what the prover uses): my rows + a halo of 5 rows of every column, nothing else needed
        part = lde.clone()
        parallel.share_row_ranges(part, world, rank, halo=5)
        lo, hi = parallel.row_range(N, rank, world)
        need = [(lo + k) % N for k in range(hi - lo + 5)]
        want_full = torch.from_numpy(oracle.lde(cols_np, 1).view(np.int64))
        assert torch.equal(part[:, need], want_full[:, need]), "row-range exchange"
        parallel.share_columns(lde, world)
        sub = np.ascontiguousarray(lde[:, lo:hi].numpy().view(np.uint64))
        my_root = oracle.merkle_build(oracle.TREE_KECCAK_M20, sub)[2]
        roots = parallel.gather_subroots(my_root, world, "cpu")
        q.put((rank, lde.numpy().view(np.uint64).copy(), roots))
    finally:
        dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process(oracle):
    rng = np.random.default_rng(11)
    cols = oracle.random_felts(rng, 5, 1 << 6)
    want_lde = oracle.lde(cols, 1)
    _, _, want_root = oracle.merkle_build(oracle.TREE_KECCAK_M20, want_lde)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cols, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=60) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, lde, roots in results:
        assert np.array_equal(lde, want_lde), f"rank {rank}: exchanged LDE differs"
        assert len(roots) == 2
        combined = oracle.hash_bytes(oracle.HASH_KECCAK_M20, roots[0] + roots[1])
        assert combined == want_root


def _ranges_worker(rank, world, port, full_np, halo, q):
    from sandstorm_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.from_numpy(full_np)
        n_cols, N = full.shape[0], full.shape[1]
        mine = torch.zeros_like(full)
        for j in parallel.owned_columns(n_cols, rank, world):
            mine[j] = full[j]                                  # a column is complete on its owner only
        parallel.share_row_ranges(mine, world, rank, halo)
        lo, hi = parallel.row_range(N, rank, world)
        need = [(lo + k) % N for k in range(hi - lo + halo)]
        q.put((rank, bool(torch.equal(mine[:, need], full[:, need]))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,halo", [(4, 7), (4, 0), (2, 40)])
def test_row_range_exchange(world, halo):
    """share_row_ranges with 4 ranks / wrap-around halo / a halo larger than a rank's range (falls back to whole columns)."""
    rng = np.random.default_rng(world + halo)
    full = rng.integers(0, 2**62, size=(5, 64, 4), dtype=np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ranges_worker, args=(r, world, port, full, halo, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in results), results


def _redistribute_worker(rank, world, port, full_np, chunk, q):
    from sandstorm_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = torch.from_numpy(full_np)
        N = full.shape[0]
        pos = torch.arange(N)
        contig = (pos // (N // world)) == rank
        cyclic = ((pos // chunk) % world) == rank
        mine = torch.where(contig[:, None], full, torch.full_like(full, -1))      # -1 marks positions this rank does not own
        parallel.redistribute(mine, world, rank, chunk, to_cyclic=True)
        ok1 = bool(torch.equal(mine[cyclic], full[cyclic]))
        mine[~cyclic] = -1                                                         # forget what is no longer owned
        parallel.redistribute(mine, world, rank, chunk, to_cyclic=False)
        ok2 = bool(torch.equal(mine[contig], full[contig]))
        q.put((rank, ok1, ok2))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,chunk", [(2, 4), (4, 2)])
def test_four_step_ownership_exchange(world, chunk):
    """redistribute(): contiguous -> chunk-cyclic -> contiguous ownership of a column (the all-to-alls of the planned
    row-sharded NTT), every owned position arrives and nothing else is needed."""
    rng = np.random.default_rng(world * 10 + chunk)
    full = rng.integers(0, 2**62, size=(256, 4), dtype=np.int64)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_redistribute_worker, args=(r, world, port, full, chunk, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=90) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(a and b for _, a, b in results), results


def test_ownership_helpers():
    from sandstorm_b200 import parallel

    assert parallel.owned_columns(9, 0, 8) == [0, 8] and parallel.owned_columns(9, 3, 8) == [3]
    assert sorted(sum((parallel.owned_columns(10, r, 4) for r in range(4)), [])) == list(range(10))
    assert parallel.row_range(1 << 10, 3, 4) == (768, 1024)
    with pytest.raises(ValueError):
        parallel.row_range(1 << 10, 0, 3)


# ---- row-sharded transforms (parallel.ShardedTransforms) with the local kernels replaced by big-int arithmetic -------------
class BigIntShardOps:
    """CPU stand-in for parallel.DeviceShardOps with the SAME contract as the C ABI entry points ss_shard_dft /
    ss_ntt_shard (include/sandstorm_b200.h), on canonical Montgomery tensors [len, 4]."""
    P = 2**251 + 17 * 2**192 + 1
    R = 2**256

    def _get(self, t, off, count):
        a = t.numpy().view(np.uint64)
        rinv = pow(self.R, -1, self.P)
        return [(int(a[off + i][0]) | int(a[off + i][1]) << 64 | int(a[off + i][2]) << 128 | int(a[off + i][3]) << 192) * rinv % self.P for i in range(count)]

    def _put(self, t, off, vals):
        a = t.numpy().view(np.uint64)
        for i, v in enumerate(vals):
            m = v % self.P * self.R % self.P
            a[off + i] = [(m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]

    def dft(self, src, src_off, src_stride, dst, dst_off, dst_stride, count, log_w, inverse, tw_log_m, tw_offset):
        P, W = self.P, 1 << log_w
        wW = pow(3, (P - 1) >> log_w, P)
        base = pow(3, (P - 1) >> tw_log_m, P) if tw_log_m >= 0 else 1
        if inverse:
            wW, base = pow(wW, -1, P), pow(base, -1, P)
        x = [self._get(src, src_off + j * src_stride, count) for j in range(W)]
        for k1 in range(W):
            out = []
            for i in range(count):
                v = sum(x[j][i] * pow(wW, j * k1, P) for j in range(W))
                out.append(v * pow(base, (tw_offset + i) * k1, P) % P)
            self._put(dst, dst_off + k1 * dst_stride, out)

    @staticmethod
    def _brev(v, bits):
        return int(f"{v:0{bits}b}"[::-1], 2) if bits else 0

    def ntt_shard(self, src, log_m, stages, log_expand, c0, h0, tw, dst):
        P, m = self.P, 1 << log_m
        vals = self._get(src, 0, m)
        if stages & 1:                                   # inverse DIF: natural -> bit-reversed, coefficient k scaled by c0 h0^k
            wi = pow(3, -((P - 1) >> log_m), P)
            coef = [sum(vals[j] * pow(wi, j * k, P) for j in range(m)) * c0 * pow(h0, k, P) % P for k in range(m)]
            vals = [coef[self._brev(p, log_m)] for p in range(m)]
        if stages & 2:                                   # forward DIT from bit-reversed coefficients, zero-padded
            mN = m << log_expand
            w = pow(3, (P - 1) >> (log_m + log_expand), P)
            coef = [vals[self._brev(k, log_m)] for k in range(m)]
            out = [sum(coef[j] * pow(w, j * k, P) for j in range(m)) * (pow(tw, k, P) if tw is not None else 1) % P for k in range(mN)]
            vals = out
        self._put(dst, 0, vals)


def _shard_worker(rank, world, port, trace_np, comp_np, log_n, log_b, q):
    from sandstorm_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        st = parallel.ShardedTransforms(rank, world, BigIntShardOps(), "cpu")
        n, N = 1 << log_n, 1 << (log_n + log_b)
        # this rank only sees its pieces of the inputs (everything else is poisoned)
        def only_mine(full, log_len):
            t = torch.full_like(torch.from_numpy(full.view(np.int64)), -1)
            for lo, cnt in parallel.pieces(log_len, rank, world):
                t[lo:lo + cnt] = torch.from_numpy(full.view(np.int64))[lo:lo + cnt]
            return t
        out = {}
        dst = torch.full((N, 4), -1, dtype=torch.int64)
        st.lde(only_mine(trace_np, log_n), log_n, log_b, dst)
        out["lde"] = dst.numpy().view(np.uint64).copy()
        dst2 = torch.full((N, 4), -1, dtype=torch.int64)
        st.lde(only_mine(trace_np, log_n), log_n, log_b, dst2, src_on_coset=True)
        out["lde_coset"] = dst2.numpy().view(np.uint64).copy()
        cols = [torch.full((N, 4), -1, dtype=torch.int64) for _ in range(2)]
        shares = st.composition_columns(only_mine(comp_np, log_n + log_b), log_n, log_b, cols)
        out["comp"] = [c.numpy().view(np.uint64).copy() for c in cols]
        out["shares"] = [s.numpy().view(np.uint64).copy() for s in shares]
        q.put((rank, out))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n,log_b", [(2, 4, 1), (4, 5, 1), (2, 4, 2)])
def test_row_sharded_transforms_match_the_single_process_definitions(oracle, world, log_n, log_b):
    """LDE of a trace column, extension of a coset-evaluated column (the DEEP quotient) and the composition-column split,
    each sharded over `world` gloo ranks with ONE all-to-all per transform side, against the oracle's whole-vector
    results: every rank must end up with exactly its block-cyclic pieces."""
    from sandstorm_b200 import parallel

    P = oracle.P
    n, N = 1 << log_n, 1 << (log_n + log_b)
    rng = np.random.default_rng(world * 100 + log_n)
    def coset_interp(evals_mont):                                  # evaluations on 3<w> -> plain coefficients (ints)
        c = oracle.from_mont(oracle.ntt(evals_mont[None], inverse=True)[0])
        return [v * pow(3, -k, P) % P for k, v in enumerate(c)]

    def coset_eval(coeffs, size):                                  # plain coefficients -> evaluations on 3<w_size> (Montgomery)
        a = [c * pow(3, k, P) % P for k, c in enumerate(coeffs)] + [0] * (size - len(coeffs))
        return oracle.ntt(oracle.to_mont(a)[None])[0]

    trace = oracle.random_felts(rng, 1, n)[0]
    comp = oracle.random_felts(rng, 1, N)[0]
    if log_b > 1:                                             # a composition polynomial of degree < 2n, as a valid trace gives
        comp = coset_eval(coset_interp(comp)[:2 * n], N)
    want_lde = oracle.lde(trace[None], log_b)[0]
    want_coset = coset_eval(coset_interp(trace), N)           # a column already on 3<w_n>, extended to 3<w_N>
    coeffs = coset_interp(comp)
    want_comp = [coset_eval(coeffs[e:2 * n:2], N) for e in range(2)]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_shard_worker, args=(r, world, port, trace, comp, log_n, log_b, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    poison = np.full(4, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
    for rank, out in results.items():
        mine = np.zeros(N, dtype=bool)
        for lo, cnt in parallel.pieces(log_n + log_b, rank, world):
            mine[lo:lo + cnt] = True
        for name, want in (("lde", want_lde), ("lde_coset", want_coset)):
            assert np.array_equal(out[name][mine], want[mine]), (name, rank)
            assert (out[name][~mine] == poison).all(), f"{name}: rank {rank} wrote rows it does not own"
        for e in range(2):
            assert np.array_equal(out["comp"][e][mine], want_comp[e][mine]), ("composition column", e, rank)
        # the returned coefficient shares: coefficient i = rank + W j2 of column e, scaled by 3^i, at position brev(j2)
        m = n // world
        bits = m.bit_length() - 1
        for e in range(2):
            got = oracle.from_mont(out["shares"][e])
            for j2 in range(m):
                i = rank + world * j2
                assert got[int(f"{j2:0{bits}b}"[::-1], 2) if bits else 0] == coeffs[2 * i + e] * pow(3, i, P) % P


# ---- the transform pairs of the OOD / DEEP stages (prover._pole_sums_on_coset and the tap-heavy OOD columns), sharded ---------------
def _pole_worker(rank, world, port, log_n, weights, z, trace_np, q):
    from sandstorm_b200 import parallel

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        P = BigIntShardOps.P
        ops = BigIntShardOps()
        st = parallel.ShardedTransforms(rank, world, ops, "cpu")
        n = 1 << log_n
        g = pow(3, (P - 1) // n, P)
        log_w = world.bit_length() - 1
        # (1) pole sums: sparse weights w_off g^-off, c0 = C z^(n-1), h0 = 3/z, no expansion  (prover._pole_sums_on_coset, world > 1)
        K, zn = pow(3, n, P), pow(z, n, P)
        c0 = pow((K - zn) % P, -1, P) * pow(z, n - 1, P) % P
        h0 = 3 * pow(z, -1, P) % P
        buf = torch.full((n, 4), -1, dtype=torch.int64)
        for lo, cnt in parallel.pieces(log_n, rank, world):
            buf[lo:lo + cnt].zero_()
        acc = {}
        for off, wgt in weights.items():
            acc[off % n] = (acc.get(off % n, 0) + wgt * pow(g, -off, P)) % P
        for o, v in acc.items():
            ops._put(buf, o, [v])                        # (also outside the owned pieces: never read there)
        share = st.to_coefficients(buf, log_n, c0 * pow(h0, rank, P) % P, pow(h0, world, P))
        out = torch.full((n, 4), -1, dtype=torch.int64)
        st.from_coefficients(share, log_n - log_w, 0, out)
        # (2) out-of-domain values of a tap-heavy column: T(z g^j) for every j  (prover._prove_sharded, heavy OOD columns)
        t = torch.full((n, 4), -1, dtype=torch.int64)
        full = torch.from_numpy(trace_np.view(np.int64))
        for lo, cnt in parallel.pieces(log_n, rank, world):
            t[lo:lo + cnt] = full[lo:lo + cnt]
        share = st.to_coefficients(t, log_n, pow(n, -1, P) * pow(z, rank, P) % P, pow(z, world, P))
        on_z = torch.full((n, 4), -1, dtype=torch.int64)
        st.from_coefficients(share, log_n - log_w, 0, on_z)
        q.put((rank, out.numpy().view(np.uint64).copy(), on_z.numpy().view(np.uint64).copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,log_n", [(2, 4), (4, 5)])
def test_sharded_pole_sums_and_coset_evaluation(oracle, world, log_n):
    """The two transform pairs this round added to the sharded prover, over gloo with big-int local ops: every rank's owned
    pieces of  sum_off w_off / (3 g^i - z g^off)  and of  T(z g^i)  equal the definitions."""
    import random

    from sandstorm_b200 import parallel

    P = oracle.P
    n = 1 << log_n
    g = pow(3, (P - 1) // n, P)
    rnd = random.Random(world * 10 + log_n)
    z = rnd.randrange(P)
    weights = {off: rnd.randrange(P) for off in (0, 1, 3, n - 1, n + 2)}
    trace = oracle.random_felts(np.random.default_rng(log_n), 1, n)[0]
    coeffs = oracle.from_mont(oracle.ntt(trace[None], inverse=True)[0])
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_pole_worker, args=(r, world, port, log_n, weights, z, trace, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = {r: (a, b) for r, a, b in (q.get(timeout=120) for _ in procs)}
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, (poles, on_z) in results.items():
        for lo, cnt in parallel.pieces(log_n, rank, world):
            got_p, got_z = oracle.from_mont(poles[lo:lo + cnt]), oracle.from_mont(on_z[lo:lo + cnt])
            for k in range(cnt):
                i = lo + k
                x = 3 * pow(g, i, P) % P
                assert got_p[k] == sum(wgt * pow((x - z * pow(g, off, P)) % P, -1, P) for off, wgt in weights.items()) % P, (rank, i)
                pt = z * pow(g, i, P) % P
                assert got_z[k] == sum(c * pow(pt, e, P) for e, c in enumerate(coeffs)) % P, (rank, i)
