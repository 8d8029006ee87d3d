#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one process per GPU): the sharded hot path (world > 1) must produce
the same commitments, out-of-domain values, FRI roots, remainder and query openings (rows and authentication paths of
every trace tree and FRI layer) as the single-GPU path on the same seeded trace.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/check_multi_gpu.py [layout] [log_n] [keccak_m20|friendly] [capi]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402
from sandstorm_b200.prover import HotPathProver, ProofOptions  # noqa: E402


def main():
    layout = sys.argv[1] if len(sys.argv) > 1 else "starknet"
    log_n = int(sys.argv[2]) if len(sys.argv) > 2 else 18
    tree = sys.argv[3] if len(sys.argv) > 3 else "keccak_m20"
    kind = ss.TREE_FRIENDLY if tree == "friendly" else ss.TREE_KECCAK_M20
    capi = len(sys.argv) > 4 and sys.argv[4] == "capi"            # collectives through csrc/dist.cu instead of torch.distributed
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)
    g = torch.Generator(device=dev).manual_seed(7)

    def rand_cols(c, rows):
        t = torch.randint(0, 2**62, (c, rows, 4), dtype=torch.int64, device=dev, generator=g)
        t[:, :, 3] &= (1 << 58) - 1
        return t

    sharded = HotPathProver(layout, log_n, ProofOptions(tree_kind=kind, capi_collectives=capi), rank=rank, world=world)
    L = sharded.layout
    base, ext = rand_cols(L.num_base_columns, 1 << log_n), rand_cols(L.num_extension_columns, 1 << log_n)
    got = sharded.prove(ss.Matrix(base), ss.Matrix(ext), queries=True, keep_openings=True)
    torch.cuda.synchronize()
    ok = True
    if rank == 0:
        single = HotPathProver(layout, log_n, ProofOptions(tree_kind=kind), rank=0, world=1)
        want = single.prove(ss.Matrix(base), ss.Matrix(ext), queries=True, keep_openings=True)
        torch.cuda.synchronize()
        checks = {"roots": got.roots == want.roots, "fri_roots": got.fri_roots == want.fri_roots, "ood_trace": got.ood_trace == want.ood_trace,
                  "ood_composition": got.ood_composition == want.ood_composition, "remainder": np.array_equal(got.remainder, want.remainder),
                  "positions": got.query_positions == want.query_positions and len(got.query_positions) > 0,
                  "trace_openings": all(np.array_equal(got.trace_queries[k][f], want.trace_queries[k][f])
                                        for k in ("base", "ext", "composition") for f in ("rows", "paths")),
                  "fri_openings": len(got.fri_layers) == len(want.fri_layers) and all(
                      a["positions"] == b["positions"] and np.array_equal(a["rows"], b["rows"]) and np.array_equal(a["paths"], b["paths"])
                      for a, b in zip(got.fri_layers, want.fri_layers)),
                  "opened_bytes": got.opened_bytes == want.opened_bytes}
        ok = all(checks.values())
        print({"layout": layout, "log_n": log_n, "world": world, "capi": capi, **checks, "ok": ok}, flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
