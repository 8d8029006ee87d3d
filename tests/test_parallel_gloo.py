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
