#!/usr/bin/env python3
"""bench.py — hot path of `sandstorm prove` on B200 (BASELINE.json metric: prove seconds & NTT
field-ops/s, 2^22-step starknet layout).

One "step" = one pass of the GPU hot path over one synthetic starknet-layout trace
(9 base + 1 extension columns of n = 16 * n_steps rows, blowup 2, 2 composition columns,
SURVEY.md §8 sizes "C3"):

    base trace   : LDE (iNTT n + coset NTT 2n per column) -> Merkle commit (masked Keccak, 9 cols)
    ext trace    : LDE -> Merkle commit (1 col, raw-leaf variant)
    composition  : [constraint evaluation when built, else a seeded stand-in column] ->
                   coset iNTT (2n) -> split into 2 columns -> coset NTT (2n) each -> Merkle commit

JSON line (driver contract):
  metric/value  = NTT field-ops/s = (1.5 N log2 N per transform, summed over the step's transforms)
                  / (device time of the step's LDE/NTT stages), inputs resident in HBM;
  ms_per_step   = whole hot-path step (all stages) = "prove seconds" * 1000, also in prove_seconds;
  e2e           = same metric through the host-buffer path (pinned host trace -> H2D -> LDE -> commit
                  -> D2H roots), copies inside the timed region;
  roofline      = ntt_pass_kernel: algorithmic bytes (2 * 32 B per element per pass) / CUDA-event time
                  of the LDE calls, against MEASURED_PEAKS.json hbm_gbs;
  cpu_baseline  = the CPU oracle (oracle/, "port") on a bounded sample, all host threads.
`--impl reference` times the CPU oracle only (the reference is Rust + un-vendored crates and cannot
be built in this image) and prints the same line with "impl": "reference".
"""
from __future__ import annotations

import argparse
import ctypes
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_BASE, N_EXT, N_COMP = 9, 1, 2            # layouts/src/starknet/air.rs:109-110 ; ce_blowup_factor = 2
LOG_BLOWUP = 1                              # cli/src/main.rs:53-54 (lde_blowup_factor = 2)
CYCLE_HEIGHT_LOG = 4                        # n = 16 * n_steps (layouts/src/starknet/mod.rs)


def workload_name(log_n: int) -> str:
    return (f"starknet layout, 2^{log_n - CYCLE_HEIGHT_LOG} Cairo steps (n=2^{log_n} rows, LDE 2^{log_n + LOG_BLOWUP}), Fp252, "
            f"{N_BASE}+{N_EXT} trace + {N_COMP} composition columns, masked-Keccak Merkle")


def ntt_ops(log_len: int) -> float:
    return 1.5 * (1 << log_len) * log_len


def lde_ops(n_cols: int, log_n: int) -> float:
    return n_cols * (ntt_ops(log_n) + ntt_ops(log_n + LOG_BLOWUP))


NTT_LOG_TILE = 11                           # csrc/ntt_fp252.cuh SS_NTT_LOG_TILE
MUL_PEAK = 592 * 1.965e9 / 275 * 32         # Fp252 multiplications/s if the IMAD pipe did nothing else (tools/ubench/bfly.cu)


def plan_passes(log_n: int) -> int:
    return 1 if log_n <= NTT_LOG_TILE else -(-log_n // NTT_LOG_TILE)


def lde_algo_bytes(n_cols: int, log_n: int) -> float:
    """2 x 32 B per element per pass; the expanding first DIT pass reads n and writes N."""
    n, N = 1 << log_n, 1 << (log_n + LOG_BLOWUP)
    pi, pf = plan_passes(log_n), plan_passes(log_n + LOG_BLOWUP)
    return n_cols * 32.0 * (2 * n * pi + (n + N) + 2 * N * (pf - 1))


def ntt_algo_bytes(n_cols: int, log_len: int) -> float:
    return n_cols * 64.0 * (1 << log_len) * plan_passes(log_len)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index), "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def __exit__(self, *exc):
        if self.proc:
            self.proc.terminate()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------ CPU arm
def cpu_oracle_run(steps: int, warmup: int, log_n: int = 18, n_cols: int = 2):
    """Bounded sample of the same workload on the host cores: LDE (+ row hashing) of n_cols columns of
    2^log_n rows with the plain-C oracle (OpenMP, all threads).  Returns (field_ops_per_s, info)."""
    import numpy as np

    import oracle

    oracle.build()
    rng = np.random.default_rng(0xB200)
    cols = oracle.random_felts(rng, n_cols, 1 << log_n)
    ops = lde_ops(n_cols, log_n)
    # torchrun pins OMP_NUM_THREADS=1; use the host's cores, and pick the thread count that is actually
    # fastest (all hardware threads vs one per physical core): the baseline should not be handicapped
    best, best_t = None, None
    for nt in sorted({os.cpu_count() or 1, max(1, (os.cpu_count() or 2) // 2), max(1, (os.cpu_count() or 4) // 4)}):
        oracle.set_threads(nt)
        t0 = time.perf_counter()
        oracle.lde(cols, LOG_BLOWUP)
        dt = time.perf_counter() - t0
        if best is None or dt < best:
            best, best_t = dt, nt
    oracle.set_threads(best_t)
    t_ntt = t_all = 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        lde = oracle.lde(cols, LOG_BLOWUP)
        t1 = time.perf_counter()
        oracle.hash_rows(oracle.HASH_KECCAK_M20, lde)
        t2 = time.perf_counter()
        t_ntt += t1 - t0
        t_all += t2 - t0
    info = {"cores": oracle.num_threads(), "sample": f"LDE of {n_cols} x 2^{log_n} Fp252 columns (blowup 2) + masked-Keccak row hashing, {steps} reps",
            "lde_s_per_rep": t_ntt / steps, "lde_plus_hash_s_per_rep": t_all / steps}
    return ops * steps / t_ntt, info


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    value, info = cpu_oracle_run(max(1, args.steps), args.warmup)
    line = {
        "impl": "reference", "metric": "ntt_field_ops_per_s", "value": value, "unit": "field-ops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": info["lde_s_per_rep"] * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "u256 (Fp252 Montgomery, 4 x u64)", "data": "synthetic",
        "config": {"workload": workload_name(args.log_steps + CYCLE_HEIGHT_LOG), "requested_log_steps": args.log_steps,
                   "sample": "bounded CPU sample of the workload's LDE stage (2 columns of 2^18 rows per step), all host threads",
                   "note": "reference binary unavailable (Rust toolchain and ministark crates absent): restated CPU oracle, OpenMP"},
        "cpu_baseline": {"value": value, "unit": "field-ops/s", "cores": info["cores"], "kind": "port", "sample": info["sample"]},
        "e2e": {"value": value, "unit": "field-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
class HotPath:
    """Device-resident buffers + the per-step stage sequence.  Columns are sharded over ranks for the
    LDE (BASELINE north_star plan A); the Merkle stage hashes this rank's row range."""

    def __init__(self, log_n: int, rank: int, world: int, seed: int = 0xB200):
        import torch

        import sandstorm_b200 as ss
        from sandstorm_b200.merkle import MatrixMerkleTree

        self.torch, self.ss, self.Tree = torch, ss, MatrixMerkleTree
        self.log_n, self.log_N = log_n, log_n + LOG_BLOWUP
        self.n, self.N = 1 << log_n, 1 << (log_n + LOG_BLOWUP)
        self.rank, self.world = rank, world
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed)

        def rand_cols(c, rows):
            t = torch.randint(0, 2**62, (c, rows, 4), dtype=torch.int64, device=dev, generator=g)
            t[:, :, 3] &= (1 << 58) - 1          # < 2^250 < p : canonical Montgomery residues
            return t

        self.base = rand_cols(N_BASE, self.n)
        self.ext = rand_cols(N_EXT, self.n)
        self.comp_evals = rand_cols(1, self.N)   # stand-in for the constraint-evaluation output
        self.base_lde = torch.empty((N_BASE, self.N, 4), dtype=torch.int64, device=dev)
        self.ext_lde = torch.empty((N_EXT, self.N, 4), dtype=torch.int64, device=dev)
        self.comp_work = torch.empty((1, self.N, 4), dtype=torch.int64, device=dev)
        self.comp_lde = torch.empty((N_COMP, self.N, 4), dtype=torch.int64, device=dev)
        self.ctx = ss.default_context()
        self.events = []
        self.roots = {}

    # -- helpers ---------------------------------------------------------------------------------
    def _lde(self, src, dst, cols):
        c, lib, ss = self.ctx, self.ctx.lib, self.ss
        for j in cols:                               # one call per owned column keeps sharding simple
            c.check(lib.ss_lde(c.handle, ss.FIELD_FP252, ctypes.c_void_p(src[j].data_ptr()), self.n, 1, self.log_n, LOG_BLOWUP,
                               ctypes.c_void_p(dst[j].data_ptr()), self.N, None, 0, ss.ORDER_NATURAL, None))

    def _owned(self, n_cols):
        from sandstorm_b200.parallel import owned_columns

        return owned_columns(n_cols, self.rank, self.world)

    def _share(self, buf, n_cols):
        from sandstorm_b200.parallel import share_columns

        share_columns(buf, self.world)               # NCCL broadcast of each LDE column from its owner

    def _commit(self, lde, kind):
        """Merkle over this rank's row range (whole matrix at world == 1)."""
        ss = self.ss
        from sandstorm_b200.parallel import row_range

        lo, hi = row_range(self.N, self.rank, self.world)
        rows = hi - lo
        sub = lde[:, lo:hi]
        c = self.ctx
        handle = ctypes.c_void_p()
        c.check(c.lib.ss_merkle_build(c.handle, kind, 0, ctypes.c_void_p(sub.data_ptr()), self.N, lde.shape[0],
                                      rows.bit_length() - 1, ss.ORDER_NATURAL, ctypes.byref(handle), None))
        return ShardTree(self, handle, kind)

    def mark(self, name):
        ev = self.torch.cuda.Event(enable_timing=True)
        ev.record()
        self.events.append((name, ev))

    # -- one step, inputs resident in HBM ------------------------------------------------------------
    def step(self):
        ss, torch = self.ss, self.torch
        self.mark("start")
        self._lde(self.base, self.base_lde, self._owned(N_BASE))
        self.mark("lde_base")
        self._share(self.base_lde, N_BASE)
        self.mark("share_base")
        t1 = self._commit(self.base_lde, ss.TREE_KECCAK_M20)
        self.mark("merkle_base")
        self._lde(self.ext, self.ext_lde, self._owned(N_EXT))
        self.mark("lde_ext")
        self._share(self.ext_lde, N_EXT)
        self.mark("share_ext")
        t2 = self._commit(self.ext_lde, ss.TREE_KECCAK_M20)
        self.mark("merkle_ext")
        # composition: evaluations on the LDE coset -> coefficients -> 2 interleaved columns -> LDE coset
        # (2 columns: done redundantly on every rank — column sharding cannot balance this phase, §8e)
        self.comp_work.copy_(self.comp_evals)
        m = ss.Matrix(self.comp_work, self.ctx)
        self.mark("comp_copy")
        m.ntt_(inverse=True, coset=True)
        self.mark("ntt_comp_inv")
        coeffs = self.comp_work.view(self.n, N_COMP, 4)
        self.comp_lde.zero_()
        self.comp_lde[:, :self.n] = coeffs.permute(1, 0, 2)
        self.mark("comp_split")
        ss.Matrix(self.comp_lde, self.ctx).ntt_(coset=True)
        self.mark("ntt_comp_fwd")
        t3 = self._commit(self.comp_lde, ss.TREE_KECCAK_M20)
        self.mark("merkle_comp")
        return t1, t2, t3

    def free(self, trees):
        for t in trees:
            t.free()

    NTT_STAGES = ("lde_base", "lde_ext", "ntt_comp_inv", "ntt_comp_fwd")

    def ntt_field_ops(self):
        return lde_ops(N_BASE, self.log_n) + lde_ops(N_EXT, self.log_n) + ntt_ops(self.log_N) + N_COMP * ntt_ops(self.log_N)

    def ntt_algo_bytes(self):
        return (lde_algo_bytes(N_BASE, self.log_n) + lde_algo_bytes(N_EXT, self.log_n) + ntt_algo_bytes(1, self.log_N) + ntt_algo_bytes(N_COMP, self.log_N))

    def ntt_launches(self):
        per_lde = plan_passes(self.log_n) + plan_passes(self.log_N)
        return (len(self._owned(N_BASE)) + len(self._owned(N_EXT))) * per_lde + 2 * plan_passes(self.log_N)


class FullHotPath(HotPath):
    """Every device stage of the prove loop (sandstorm_b200/prover.py) with the real starknet AIR, on 1..8 ranks."""

    NTT_STAGES = ("lde_base", "lde_ext", "ntt_comp_inv", "ntt_comp_fwd", "deep_lde")

    def ntt_field_ops(self):
        return HotPath.ntt_field_ops(self) + ntt_ops(self.log_n) + ntt_ops(self.log_N)      # + extension of the DEEP quotient

    def ntt_algo_bytes(self):
        return HotPath.ntt_algo_bytes(self) + ntt_algo_bytes(1, self.log_n) + ntt_algo_bytes(1, self.log_N)

    def __init__(self, log_n: int, rank: int = 0, world: int = 1, seed: int = 0xB200):
        import torch

        import sandstorm_b200 as ss
        from sandstorm_b200.prover import HotPathProver

        self.torch, self.ss = torch, ss
        self.log_n, self.log_N = log_n, log_n + LOG_BLOWUP
        self.n, self.N = 1 << log_n, 1 << (log_n + LOG_BLOWUP)
        self.rank, self.world = rank, world
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed)

        def rand_cols(c, rows):
            t = torch.randint(0, 2**62, (c, rows, 4), dtype=torch.int64, device=dev, generator=g)
            t[:, :, 3] &= (1 << 58) - 1
            return t

        self.base, self.ext = rand_cols(N_BASE, self.n), rand_cols(N_EXT, self.n)
        self.ctx = ss.default_context()
        from sandstorm_b200.prover import ProofOptions

        self.prover = HotPathProver("starknet", log_n, ProofOptions(col_pad_rows=int(os.environ.get("SS_COL_PAD_ROWS", "0"))), rank=rank, world=world)
        t0 = time.perf_counter()
        # challenge-independent structure pass (per layout and trace length); the per-proof value patch runs INSIDE the step
        self.prover.composition_template()
        self.compile_s = time.perf_counter() - t0
        self.events = self.prover.timeline
        self.last = None
        self.column_ready = None

    def step(self):
        ss = self.ss
        self.last = self.prover.prove(ss.Matrix(self.base, self.ctx), ss.Matrix(self.ext, self.ctx), column_ready=self.column_ready)
        return []

    def free(self, trees):
        pass

    def ntt_launches(self):
        return 0


class ShardTree:
    """This rank's sub-tree of a row-sharded commitment; root() all-gathers the sub-roots (32 B per
    rank over NCCL) and combines them with ss_merkle_combine."""

    def __init__(self, hp, handle, kind):
        self.hp, self.handle, self.kind = hp, handle, kind

    def root(self) -> bytes:
        hp = self.hp
        c, torch = hp.ctx, hp.torch
        out = (ctypes.c_uint8 * 32)()
        c.check(c.lib.ss_merkle_root(c.handle, self.handle, out))
        if hp.world == 1:
            return bytes(out)
        from sandstorm_b200.parallel import gather_subroots

        roots = gather_subroots(bytes(out), hp.world, "cuda")
        sub = (ctypes.c_uint8 * (32 * hp.world)).from_buffer_copy(b"".join(roots))
        c.check(c.lib.ss_merkle_combine(c.handle, self.kind, sub, hp.world.bit_length() - 1, out))
        return bytes(out)

    def free(self):
        if self.handle:
            self.hp.ctx.lib.ss_tree_free(self.handle)
            self.handle = None


def stage_times(events):
    out = {}
    for (_, a), (name, b) in zip(events[:-1], events[1:]):
        if name != "start":
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return out


def gpu_arm(args):
    import torch

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; sandstorm_b200 has no CPU fallback (use --impl reference for the CPU oracle)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import sandstorm_b200 as ss

    log_n = args.log_steps + CYCLE_HEIGHT_LOG
    free_b, _ = torch.cuda.mem_get_info()
    need = lambda ln: 32.0 * ((N_BASE + N_EXT) * (1 << ln) * 2 + (N_BASE + N_EXT + N_COMP + 2) * (2 << ln) + 3 * 2 * (2 << ln))
    while need(log_n) > 0.85 * free_b and log_n > 12:
        log_n -= 1
    need_full = lambda ln: 32.0 * ((N_BASE + N_EXT) * (1 << ln) * 2 + (N_BASE + N_EXT + N_COMP + 3) * (2 << ln) + 4 * (2 << ln) + 3 * 2 * (2 << ln)) + 6e9
    if not args.partial:
        while need_full(log_n) > 0.9 * free_b and log_n > 15:
            log_n -= 1
        hp = FullHotPath(log_n, rank, world)
    else:
        hp = HotPath(log_n, rank, world)
    ctx = hp.ctx

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(args.warmup):
        hp.free(hp.step())
    barrier()
    hp.events.clear()
    l0 = ctx.lib.ss_kernel_launches(ctx.handle)
    with ClockSampler(local) as clk:
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        per_step = []
        for _ in range(args.steps):
            n0 = len(hp.events)
            hp.free(hp.step())
            per_step.append((n0, len(hp.events)))
        t_end.record()
        barrier()
    clocks = clk.summary()
    launches = (ctx.lib.ss_kernel_launches(ctx.handle) - l0) / args.steps
    total_ms = t_start.elapsed_time(t_end)
    stages = {}
    for a, b in per_step:
        for k, v in stage_times(hp.events[a:b]).items():
            stages[k] = stages.get(k, 0.0) + v / args.steps
    ntt_ms = sum(stages.get(k, 0.0) for k in type(hp).NTT_STAGES)
    t = torch.tensor([total_ms / args.steps, ntt_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step, ntt_ms = float(t[0]), float(t[1])
    value = hp.ntt_field_ops() / (ntt_ms * 1e-3)

    # ---- end-to-end through host buffers (pinned host trace -> H2D -> LDE -> commit -> D2H roots) ---
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        nb_cols = hp.base.shape[0]
        n_trace_cols = nb_cols + hp.ext.shape[0]
        dev_col = lambda k: hp.base[k] if k < nb_cols else hp.ext[k - nb_cols]
        # what this rank has to upload: whole columns it transforms, and of the others only the rows it reads
        owned, pieces = None, [(0, hp.n)]
        if isinstance(hp, FullHotPath) and world > 1:
            owned = set(hp.prover.trace_columns_owned())
            lo, cnt = hp.prover.trace_rows_needed()
            pieces = [(lo, min(hp.n, lo + cnt))] + ([(0, lo + cnt - hp.n)] if lo + cnt > hp.n else [])
        # pinned host copy of exactly those rows, per column: [(device slice, pinned host tensor), ...]
        plan = []
        for k in range(n_trace_cols):
            parts = [(0, hp.n)] if (owned is None or k in owned) else pieces
            plan.append([(dev_col(k)[a:b], dev_col(k)[a:b].cpu().pin_memory()) for a, b in parts])
        my_h2d = sum(h.numel() * 8 for col in plan for _, h in col)
        def e2e_step():
            # the trace is uploaded column by column on a copy stream; the LDE of column k waits only for column k,
            # so the rest of the H2D traffic overlaps the first LDE stage
            main = torch.cuda.current_stream()
            copy_stream.wait_stream(main)                  # the previous step has finished reading the buffers
            events = []
            with torch.cuda.stream(copy_stream):
                for k in range(n_trace_cols):
                    for dst, src in plan[k]:
                        dst.copy_(src, non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_stream)
                    events.append(ev)
            if isinstance(hp, FullHotPath):
                hp.column_ready = lambda k: main.wait_stream(copy_stream) if k is None else main.wait_event(events[k])
            else:
                main.wait_stream(copy_stream)
            trees = hp.step()
            main.wait_stream(copy_stream)
            roots = [tr.root() for tr in trees]    # D2H of each 32-byte root (+ sub-root all-gather at N > 1)
            hp.free(trees)                         # (the full prover reads its roots, OOD values and openings itself)
            return roots

        for _ in range(max(1, args.warmup - 2)):
            e2e_step()
        barrier()
        hp.events.clear()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        e2e_ms = e0.elapsed_time(e1) / args.steps
        t = torch.tensor([e2e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_ms = float(t[0])
        hb = torch.tensor([float(my_h2d)], dtype=torch.float64, device="cuda")
        if world > 1:
            torch.distributed.all_reduce(hb)
        h2d = int(hb[0])                               # summed over the ranks
        d2h = 32 * 3
        if getattr(hp, "last", None) is not None:
            r = hp.last
            d2h = 32 * (3 + len(r.fri_roots)) + 32 * (len(r.ood_trace) + len(r.ood_composition)) + r.remainder.nbytes + r.opened_bytes
        e2e = {"value": hp.ntt_field_ops() / (e2e_ms * 1e-3), "unit": "field-ops/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
               "note": "whole committed-LDE call incl. copies, hashing and tree build in the denominator"}
        del plan

    if rank != 0:
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    achieved = hp.ntt_algo_bytes() / world / (ntt_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "ntt_pass_traffic.json")
    if os.path.exists(tp):
        traffic = json.load(open(tp)).get("dram_bytes_per_launch")
    # dominant kernel of the step: the composition-constraint evaluation (one launch per step; its stage is the launch)
    roofline = None
    if isinstance(hp, FullHotPath) and stages.get("constraint_eval"):
        ce_ms = stages["constraint_eval"]
        n_cols_read = N_BASE + N_EXT + 1                                   # trace columns + the w = 1/(x-1) column
        rows = hp.N // world
        algo = 32.0 * (n_cols_read + 1) * rows                              # every column element once + one output per row
        ce_ach = algo / (ce_ms * 1e-3) / 1e9
        ce_traffic = None
        cp = os.path.join(ROOT, "profiles", "ce_kernel_traffic.json")
        if os.path.exists(cp):
            t = json.load(open(cp))
            ce_traffic = t["dram_bytes_per_row"] * rows
        prog = hp.prover._composition_program
        muls = prog.n_mul * rows / (ce_ms * 1e-3)
        roofline = {"bound": "hbm", "kernel": "ce_gen_starknet_composition (ss_constraint_eval)", "achieved": ce_ach, "peak": peak, "unit": "GB/s",
                    "frac": ce_ach / peak, "traffic": ce_traffic, "peak_source": peak_src, "launch_ms": ce_ms,
                    "algorithmic_bytes_per_row": 32 * (n_cols_read + 1),
                    "field_muls_per_s": muls, "field_mul_pipe_peak_per_s": MUL_PEAK, "field_mul_pipe_frac": muls / MUL_PEAK,
                    "note": "arithmetic-bound: %d Montgomery multiplications + %d add/sub per row on 384 algorithmic bytes; the integer pipe, not HBM, is the roof "
                            "(mul peak = 592 SMSPs x 1.965 GHz / 275 cycles per warp-multiplication, profiles/r01_ntt_tile_ab.md)" % (prog.n_mul, prog.n_addsub)}
    cpu_val, cpu_info = (None, {})
    if not args.no_cpu and world >= 1:
        cpu_val, cpu_info = cpu_oracle_run(2, 1)
    line = {
        "metric": "ntt_field_ops_per_s", "value": value, "unit": "field-ops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "prove_seconds": ms_per_step / 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u256 (Fp252 Montgomery, 8 x u32 limbs)", "data": "synthetic",
        "config": {"workload": workload_name(log_n),
                   "requested_log_steps": args.log_steps, "parallelism": f"{world} rank(s): LDE sharded by column (NCCL broadcast), Merkle + constraint eval + DEEP + FRI folds by LDE row range and OOD by trace row range (all-gather, combined sub-roots / summed partial values); composition-column and DEEP-extension NTTs replicated", "l2": "inputs_larger_than_L2",
                   "stages_in_step": list(stages.keys()),
                   "not_in_step": [] if isinstance(hp, FullHotPath) else ["constraint_eval (stand-in column)", "ood", "deep_composition", "fri_layers", "queries"],
                   "air": "starknet layout, 195 constraints (sandstorm_b200/air/layouts/starknet.json)" if isinstance(hp, FullHotPath) else None,
                   "template_compile_s_outside_step": round(getattr(hp, "compile_s", 0.0), 1)},
        "stages_ms": {k: round(v, 3) for k, v in stages.items()},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_ntt": {"bound": "hbm", "kernel": "ss::ntt_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": traffic, "peak_source": peak_src,
                         "note": "Fp252 NTT is bound by the carry-chained IMAD.WIDE pipe, not HBM (profiles/r01_pipe_microbench.md)"},
        "cpu_baseline": {"value": cpu_val, "unit": "field-ops/s", "cores": cpu_info.get("cores"), "kind": "port", "sample": cpu_info.get("sample")},
    }
    if e2e:
        line["e2e"] = e2e
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--log-steps", type=int, default=22, help="log2 of Cairo steps (22 = BASELINE metric config; n = 16 * steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--partial", action="store_true", help="LDE + commits only, with a stand-in composition column")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    else:
        gpu_arm(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
