#!/usr/bin/env python3
"""NTT microbench (BASELINE.json configs[4]): batched Fp252 forward DIF + inverse DIT, 2^16..2^26,
field-ops/s (1.5 N log2 N per transform) and algorithmic HBM GB/s (2 * N * 32 B per transform)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sandstorm_b200 as ss  # noqa: E402


def main():
    logs = [int(a) for a in sys.argv[1:]] or [16, 18, 20, 22, 24]
    torch.cuda.set_device(0)
    out = []
    for log_n in logs:
        n = 1 << log_n
        n_cols = max(1, min(64, (1 << 27) // n))           # ~4 GiB per batch
        g = torch.Generator(device="cuda").manual_seed(0xB200 + log_n)
        data = torch.randint(0, 2**62, (n_cols, n, 4), dtype=torch.int64, device="cuda", generator=g)
        data[:, :, 3] &= (1 << 58) - 1                     # < 2^250 < p: canonical residues
        m = ss.Matrix(data)
        for _ in range(2):
            m.ntt_(out_order=ss.ORDER_BITREV)
            m.ntt_(inverse=True, in_order=ss.ORDER_BITREV)
        torch.cuda.synchronize()
        ref = data.clone()
        reps = 5
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        ev[0].record()
        for _ in range(reps):
            m.ntt_(out_order=ss.ORDER_BITREV)
        ev[1].record()
        for _ in range(reps):
            m.ntt_(inverse=True, in_order=ss.ORDER_BITREV)
        ev[2].record()
        torch.cuda.synchronize()
        fwd_ms = ev[0].elapsed_time(ev[1]) / reps
        inv_ms = ev[1].elapsed_time(ev[2]) / reps
        ops = 1.5 * n * log_n * n_cols
        rec = {
            "log_n": log_n, "n_cols": n_cols, "fwd_ms": round(fwd_ms, 4), "inv_ms": round(inv_ms, 4),
            "fwd_field_ops_per_s": ops / (fwd_ms * 1e-3), "inv_field_ops_per_s": ops / (inv_ms * 1e-3),
            "fwd_algo_GBps": 2 * n * 32 * n_cols / (fwd_ms * 1e-3) / 1e9,
            "per_column_fwd_us": fwd_ms * 1e3 / n_cols,
        }
        print(json.dumps(rec), flush=True)
        out.append(rec)
        del m, data, ref
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_ntt.json", "w") as f:
        json.dump(out, f, indent=1)


if __name__ == "__main__":
    main()
