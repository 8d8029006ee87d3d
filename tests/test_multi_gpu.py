"""GPU, needs >= 2 devices (skipped on a one-GPU box): the row-sharded hot path over NCCL reproduces the single-GPU
commitments, out-of-domain values, FRI roots, remainder and query openings (rows + authentication paths of every trace tree
and FRI layer) bit for bit (tools/check_multi_gpu.py under torchrun)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("layout,log_n,tree,collectives", [("recursive", 15, "friendly", "torch"), ("starknet", 19, "keccak_m20", "torch"),
                                                           ("plain", 10, "keccak_m20", "torch"), ("recursive", 15, "friendly", "capi"),
                                                           ("starknet", 19, "keccak_m20", "capi")])
def test_sharded_run_equals_single_gpu(layout, log_n, tree, collectives):
    import torch

    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs at least two CUDA devices")
    world = 4 if torch.cuda.device_count() >= 4 and log_n >= 15 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1", "--master-port", "29631",
           os.path.join(ROOT, "tools", "check_multi_gpu.py"), layout, str(log_n), tree, collectives]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-3000:]
    assert "'ok': True" in out.stdout
