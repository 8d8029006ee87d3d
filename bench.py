#!/usr/bin/env python3
"""bench.py — hot path of `sandstorm prove` on B200 (BASELINE.json metric: prove seconds & NTT field-ops/s,
2^22-step starknet layout, 1/2/4/8 B200 vs CPU).

One "step" = one pass of the whole device hot path (sandstorm_b200/prover.py) over one synthetic trace of the layout:
base LDE + commit, [challenges], extension LDE + commit, per-proof value patch of the composition program, constraint
evaluation over the LDE coset, composition columns (coset iNTT, split, coset NTT) + commit, out-of-domain values, DEEP
quotient + extension, FRI layers, query openings.  Workloads (--workload):

    starknet22  (default; BASELINE configs[2], the metric's configuration) starknet layout, 2^22 Cairo steps, masked-Keccak trees
    recursive20 (configs[1]) recursive layout, 2^20 steps, FriendlyMerkleTree<22> (Blake2s + Pedersen) commitments
    ntt         (configs[4]) batched Fp252 forward + inverse NTT microbenchmark, 2^16 .. 2^26, one batch per GPU

JSON line (driver contract):
  metric/value  = NTT field-ops/s = 1.5 N log2 N per transform, summed over the step's transforms, / device time of the step's
                  LDE/NTT stages (CUDA events, max over ranks), inputs resident in HBM;
  ms_per_step   = the whole step = prove seconds (GPU stages) * 1000;
  e2e           = the same field-op count / the whole step measured through the public API with the trace in pinned HOST
                  memory: H2D of every trace column and D2H of roots / OOD values / openings inside the timed region;
  roofline      = the dominant kernel (composition-constraint evaluation), algorithmic bytes of SURVEY §8(d) over its
                  CUDA-event time, against MEASURED_PEAKS.json; roofline_ntt = the metric's kernel on the same basis
                  (n*s + N*s per LDE column, 2*N*s per NTT) and on the 2^24-point NTT the north_star target is quoted on;
  cpu_baseline  = the CPU restatement of the SAME pipeline (oracle/prover.py, OpenMP, all host threads) on a bounded
                  sample (a shorter trace of the same layout), kind "port".
`--impl reference` runs only that CPU pipeline, every stage, and prints the same line with "impl": "reference": the
reference itself is Rust + un-vendored crates and cannot be built in this image."""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

LOG_BLOWUP = 1                              # cli/src/main.rs:53-54 (lde_blowup_factor = 2)
CE = 2                                      # air.ce_blowup_factor() for Cairo (degree-2 constraints)
CYCLE_HEIGHT_LOG = 4                        # n = 16 * n_steps (layouts/src/*/mod.rs CYCLE_HEIGHT)
WORKLOADS = {"starknet22": ("starknet", 22, "keccak_m20"), "recursive20": ("recursive", 20, "friendly")}
MUL_PEAK = 592 * 1.965e9 / 275 * 32         # Fp252 multiplications/s if the IMAD pipe did nothing else (tools/ubench/bfly.cu)
MUL_SUSTAINED = 592 * 1.965e9 / 355 * 32    # what a register-only loop of dependent fp::mul sustains with every warp timed to completion
                                            # (tools/ubench/mulmix.cu, profiles/r03_mul_variants.md): the practical ceiling of the routine


def workload_name(layout: str, log_n: int, tree: str, n_base: int, n_ext: int) -> str:
    trees = {"keccak_m20": "masked-Keccak Merkle", "friendly": "FriendlyMerkleTree<22> (Blake2s + Pedersen)"}[tree]
    return (f"{layout} layout, 2^{log_n - CYCLE_HEIGHT_LOG} Cairo steps (n=2^{log_n} rows, LDE 2^{log_n + LOG_BLOWUP}), Fp252, "
            f"{n_base}+{n_ext} trace + {CE} composition columns, {trees}")


def ntt_ops(log_len: int) -> float:
    return 1.5 * (1 << log_len) * log_len


def step_ntt_ops(n_cols: int, log_n: int) -> float:
    """field-ops of the step's transforms: LDE of every trace column, composition iNTT + ce coset NTTs, DEEP extension."""
    log_N = log_n + LOG_BLOWUP
    return n_cols * (ntt_ops(log_n) + ntt_ops(log_N)) + (1 + CE) * ntt_ops(log_N) + ntt_ops(log_n) + ntt_ops(log_N)


def step_ntt_bytes(n_cols: int, log_n: int) -> float:
    """ALGORITHMIC HBM bytes of the same transforms: n*s + N*s per LDE column, 2*N*s per NTT of size N."""
    n, N, s = 1 << log_n, 1 << (log_n + LOG_BLOWUP), 32.0
    return n_cols * (n + N) * s + (1 + CE) * 2 * N * s + (n + N) * s


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
class CpuArm:
    """The same pipeline on the host cores (oracle/prover.py), on a shorter trace of the same layout."""

    def __init__(self, layout: str, tree: str, log_n: int):
        import numpy as np

        import oracle
        from oracle.prover import CpuHotPath

        oracle.build()
        # torchrun pins OMP_NUM_THREADS=1; the baseline gets every host core
        oracle.set_threads(os.cpu_count() or 1)
        self.oracle, self.log_n = oracle, log_n
        kind = oracle.TREE_FRIENDLY if tree == "friendly" else oracle.TREE_KECCAK_M20
        self.hp = CpuHotPath(layout, log_n, log_blowup=LOG_BLOWUP, tree_kind=kind)
        L = self.hp.layout
        self.n_cols = L.num_columns
        rng = np.random.default_rng(0xB200)
        self.base = oracle.random_felts(rng, L.num_base_columns, 1 << log_n)
        self.ext = oracle.random_felts(rng, L.num_extension_columns, 1 << log_n)
        self.hp.prepare()                                    # challenge-independent template, as on the GPU arm

    def step(self):
        from sandstorm_b200.prover import SeededCoin

        t0 = time.perf_counter()
        self.hp.prove(self.base, self.ext, SeededCoin(1))
        total = time.perf_counter() - t0
        st = self.hp.stages
        ntt = sum(st.get(k, 0.0) for k in ("lde_base", "lde_ext", "ntt_comp_inv", "ntt_comp_fwd"))
        return total, ntt, dict(st)

    def run(self, steps: int, warmup: int):
        for _ in range(warmup):
            self.step()
        tot = ntt = 0.0
        stages: dict = {}
        for _ in range(steps):
            a, b, st = self.step()
            tot, ntt = tot + a, ntt + b
            for k, v in st.items():
                stages[k] = stages.get(k, 0.0) + v / steps
        # (the CPU formulation extends nothing for DEEP: it evaluates the quotient on every LDE row, like ministark)
        ops = self.n_cols * (ntt_ops(self.log_n) + ntt_ops(self.log_n + LOG_BLOWUP)) + (1 + CE) * ntt_ops(self.log_n + LOG_BLOWUP)
        return {"ntt_ops_per_s": ops * steps / ntt, "e2e_ops_per_s": ops * steps / tot, "prove_s": tot / steps, "ntt_s": ntt / steps,
                "stages_s": {k: round(v, 4) for k, v in stages.items()}, "cores": self.oracle.num_threads(),
                "sample": f"whole hot path (every stage) on a 2^{self.log_n - CYCLE_HEIGHT_LOG}-step trace of the same layout, {steps} reps"}


def reference_arm(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    layout, log_steps, tree = WORKLOADS.get(args.workload, WORKLOADS["starknet22"])
    log_steps = args.log_steps or log_steps
    sample_log_n = min(log_steps + CYCLE_HEIGHT_LOG, args.cpu_log_n)
    arm = CpuArm(layout, tree, sample_log_n)
    r = arm.run(max(1, args.steps), args.warmup)
    L = arm.hp.layout
    full_log_n = log_steps + CYCLE_HEIGHT_LOG
    scale = (1 << (full_log_n - sample_log_n)) * (full_log_n + 1) / (sample_log_n + 1)
    line = {
        "impl": "reference", "metric": "ntt_field_ops_per_s", "value": r["ntt_ops_per_s"], "unit": "field-ops/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["prove_s"] * 1e3, "prove_seconds": r["prove_s"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u256 (Fp252 Montgomery, 4 x u64)", "data": "synthetic",
        "config": {"workload": workload_name(layout, full_log_n, tree, L.num_base_columns, L.num_extension_columns), "requested_log_steps": log_steps,
                   "sample": r["sample"], "prove_seconds_extrapolated_to_workload": r["prove_s"] * scale,
                   "note": "reference binary unavailable (Rust toolchain and ministark crates absent): restated CPU pipeline, C + OpenMP, "
                           "the same stages as the GPU arm; rates (field-ops/s) are comparable across trace lengths, seconds are per sample"},
        "stages_s": r["stages_s"],
        "cpu_baseline": {"value": r["ntt_ops_per_s"], "unit": "field-ops/s", "cores": r["cores"], "kind": "port", "sample": r["sample"]},
        "e2e": {"value": r["e2e_ops_per_s"], "unit": "field-ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------ GPU arm
class FullHotPath:
    """Every device stage of the prove loop (sandstorm_b200/prover.py) with the layout's real AIR, on 1..8 ranks."""

    NTT_STAGES = ("lde_base", "lde_ext", "ntt_comp_inv", "ntt_comp_fwd", "ntt_comp", "deep_lde")

    def __init__(self, layout: str, tree: str, log_n: int, rank: int = 0, world: int = 1, seed: int = 0xB200):
        import torch

        import sandstorm_b200 as ss
        from sandstorm_b200.prover import HotPathProver, ProofOptions

        self.torch, self.ss = torch, ss
        self.log_n, self.log_N = log_n, log_n + LOG_BLOWUP
        self.n, self.N = 1 << log_n, 1 << (log_n + LOG_BLOWUP)
        self.rank, self.world = rank, world
        dev = torch.device("cuda", torch.cuda.current_device())
        g = torch.Generator(device=dev).manual_seed(seed)

        def rand_cols(c, rows):
            t = torch.randint(0, 2**62, (c, rows, 4), dtype=torch.int64, device=dev, generator=g)
            t[:, :, 3] &= (1 << 58) - 1          # < 2^250 < p : canonical Montgomery residues
            return t

        kind = ss.TREE_FRIENDLY if tree == "friendly" else ss.TREE_KECCAK_M20
        self.prover = HotPathProver(layout, log_n, ProofOptions(log_blowup=LOG_BLOWUP, tree_kind=kind, col_pad_rows=int(os.environ.get("SS_COL_PAD_ROWS", "0"))),
                                    rank=rank, world=world)
        L = self.prover.layout
        self.n_base, self.n_ext = L.num_base_columns, L.num_extension_columns
        self.base, self.ext = rand_cols(self.n_base, self.n), rand_cols(self.n_ext, self.n)
        self.ctx = ss.default_context()
        t0 = time.perf_counter()
        # challenge-independent structure pass (per layout and trace length); the per-proof value patch runs INSIDE the step
        self.prover.prepare()
        self.compile_s = time.perf_counter() - t0
        self.events = self.prover.timeline
        self.last = None
        self.column_ready = None

    def step(self):
        ss = self.ss
        self.last = self.prover.prove(ss.Matrix(self.base, self.ctx), ss.Matrix(self.ext, self.ctx), column_ready=self.column_ready)

    def ntt_field_ops(self):
        return step_ntt_ops(self.n_base + self.n_ext, self.log_n)

    def ntt_algo_bytes(self):
        return step_ntt_bytes(self.n_base + self.n_ext, self.log_n)


def stage_times(events):
    out = {}
    for (_, a), (name, b) in zip(events[:-1], events[1:]):
        if name != "start":
            out[name] = out.get(name, 0.0) + a.elapsed_time(b)
    return out


def ntt_microbench(torch, ss, log_len: int, n_cols: int, reps: int = 3):
    """forward + inverse NTT of n_cols columns of 2^log_len elements, in place (BASELINE configs[4]).  Returns
    (field-ops/s, algorithmic GB/s = 2 * N * 32 B per column per transform / time, seconds per batch transform)."""
    dev = torch.device("cuda", torch.cuda.current_device())
    t = torch.randint(0, 2**62, (n_cols, 1 << log_len, 4), dtype=torch.int64, device=dev)
    t[:, :, 3] &= (1 << 58) - 1
    m = ss.Matrix(t)
    for _ in range(2):
        m.ntt_(out_order=ss.ORDER_BITREV)
        m.ntt_(inverse=True, in_order=ss.ORDER_BITREV)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        m.ntt_(out_order=ss.ORDER_BITREV)                      # DIF: natural -> bit-reversed (no permutation pass)
        m.ntt_(inverse=True, in_order=ss.ORDER_BITREV)         # DIT: bit-reversed -> natural
    e1.record()
    torch.cuda.synchronize()
    sec = e0.elapsed_time(e1) * 1e-3 / (2 * reps)
    del m, t
    return n_cols * ntt_ops(log_len) / sec, n_cols * 2 * (1 << log_len) * 32 / sec / 1e9, sec


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def init_dist():
    import torch

    rank, local, world = int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; sandstorm_b200 has no CPU fallback (use --impl reference for the CPU pipeline)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return rank, local, world


def ntt_arm(args):
    """BASELINE configs[4]: batched forward + inverse NTT, 2^16 .. 2^26, one independent batch per GPU (weak scaling)."""
    import torch

    import sandstorm_b200 as ss

    rank, local, world = init_dist()
    peak, peak_src = peaks()
    table = {}
    with ClockSampler(local) as clk:
        for log_len in (16, 18, 20, 22, 24, 26):
            cols = max(1, min(64, (1 << 27) >> log_len))                   # ~4 GiB per batch
            ops, gbs, sec = ntt_microbench(torch, ss, log_len, cols, max(1, args.steps))
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            sec = float(t[0])
            table[f"2^{log_len}"] = {"columns_per_gpu": cols, "field_ops_per_s": world * cols * ntt_ops(log_len) / sec,
                                     "algorithmic_gbs_per_gpu": cols * 2 * (1 << log_len) * 32 / sec / 1e9, "ms_per_transform_batch": sec * 1e3}
    if rank != 0:
        return
    head = table["2^24"]
    line = {"metric": "ntt_field_ops_per_s", "value": head["field_ops_per_s"], "unit": "field-ops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": head["ms_per_transform_batch"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u256 (Fp252 Montgomery, 8 x u32 limbs)", "data": "synthetic",
            "config": {"workload": "NTT microbench: batched 2^16-2^26 Fp252 forward + inverse NTT (value: 2^24, 8 columns per GPU)", "l2": "inputs_larger_than_L2"},
            "sizes": table, "clocks": clk.summary(), "gpu_launches": 6,
            "roofline": {"bound": "hbm", "kernel": "ss::ntt_pass_kernel (2^24-point transform)", "achieved": head["algorithmic_gbs_per_gpu"], "peak": peak,
                         "unit": "GB/s", "frac": head["algorithmic_gbs_per_gpu"] / peak, "traffic": None, "peak_source": peak_src,
                         "note": "algorithmic bytes = 2 * N * 32 B per transform (read once, write once; SURVEY §8d); the kernel makes 3 passes and is bound by the carry-chained IMAD pipe"}}
    print(json.dumps(line), flush=True)


def goldilocks_arm(args):
    """BASELINE configs[3]: the single-limb NTT path on the shape of a recursive-layout trace (10 columns) over Goldilocks:
    per-column LDE (interpolate on <w_n>, evaluate on 7<w_2n>), columns split over the ranks (they shard embarrassingly:
    no exchange).  Default 2^24 Cairo steps = 2^28 rows; reduced until the columns of a rank fit its memory."""
    import torch

    from sandstorm_b200 import goldilocks as glk

    rank, local, world = init_dist()
    peak, peak_src = peaks()
    n_cols_total = 10
    log_n = (args.log_steps or 24) + CYCLE_HEIGHT_LOG
    mine = [k for k in range(n_cols_total) if k % world == rank]
    free_b, _ = torch.cuda.mem_get_info()
    while len(mine) * 8 * 4 * (1 << log_n) > 0.8 * free_b and log_n > 16:      # trace + LDE + coefficient scratch
        log_n -= 1
    g = torch.Generator(device="cuda").manual_seed(3 + rank)
    trace = torch.randint(0, 2**62, (len(mine), 1 << log_n), dtype=torch.int64, device="cuda", generator=g)

    def step():
        return glk.lde(trace, 1)

    for _ in range(args.warmup):
        step()
    if world > 1:
        torch.distributed.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            out = step()
        e1.record()
        torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / args.steps], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t[0])
    if rank != 0:
        return
    n = 1 << log_n
    ops = n_cols_total * (1.5 * n * log_n + 1.5 * 2 * n * (log_n + 1))          # inverse transform of n + forward transform of 2n per column
    per_gpu_bytes = len(mine) * (n * 8 + 2 * n * 8)                              # n s + N s per LDE column
    gbs = per_gpu_bytes / (ms * 1e-3) / 1e9
    line = {"metric": "ntt_field_ops_per_s", "value": ops / (ms * 1e-3), "unit": "field-ops/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "u64 (Goldilocks, stored words)", "data": "synthetic",
            "config": {"workload": f"Goldilocks LDE (blowup 2) of {n_cols_total} columns x 2^{log_n} rows, columns split over the ranks",
                       "requested_log_steps": args.log_steps or 24, "l2": "inputs_larger_than_L2", "columns_on_rank_0": len(mine)},
            "clocks": clk.summary(), "gpu_launches": None,
            "roofline": {"bound": "hbm", "kernel": "gl_pass_kernel (ss_lde, SS_FIELD_GOLDILOCKS)", "achieved": gbs, "peak": peak, "unit": "GB/s",
                         "frac": gbs / peak, "traffic": None, "peak_source": peak_src,
                         "note": "algorithmic bytes n*8 + N*8 per column on the busiest rank; the transform is bound by instruction issue "
                                 "(~380 instructions per element at best, DESIGN.md §4.7), not HBM"}}
    print(json.dumps(line), flush=True)


def gpu_arm(args):
    import torch

    rank, local, world = init_dist()
    import sandstorm_b200 as ss
    from sandstorm_b200.air.layouts import load_layout

    layout, log_steps, tree = WORKLOADS[args.workload]
    log_steps = args.log_steps or log_steps
    log_n = log_steps + CYCLE_HEIGHT_LOG
    L = load_layout(layout)
    C = L.num_columns
    free_b, _ = torch.cuda.mem_get_info()
    need = lambda ln: 32.0 * (C * (1 << ln) * 2 + (C + CE + 3) * (2 << ln) + 4 * (2 << ln) + 3 * 2 * (2 << ln)) + 6e9
    while need(log_n) > 0.9 * free_b and log_n > 15:
        log_n -= 1
    hp = FullHotPath(layout, tree, log_n, rank, world)
    ctx = hp.ctx

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing ----------------------------------------------------------------
    for _ in range(args.warmup):
        hp.step()
    barrier()
    hp.events.clear()
    l0 = ctx.lib.ss_kernel_launches(ctx.handle)
    with ClockSampler(local) as clk:
        t_start = torch.cuda.Event(enable_timing=True); t_end = torch.cuda.Event(enable_timing=True)
        t_start.record()
        per_step = []
        for _ in range(args.steps):
            n0 = len(hp.events)
            hp.step()
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
    ntt_ms = sum(stages.get(k, 0.0) for k in FullHotPath.NTT_STAGES)
    t = torch.tensor([total_ms / args.steps, ntt_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms_per_step, ntt_ms = float(t[0]), float(t[1])
    value = hp.ntt_field_ops() / (ntt_ms * 1e-3)

    # ---- end-to-end through host buffers: pinned host trace -> H2D -> every stage -> D2H of roots / OOD values / openings ---
    e2e = None
    if not args.no_e2e:
        copy_stream = torch.cuda.Stream()
        n_trace_cols = hp.n_base + hp.n_ext
        dev_col = lambda k: hp.base[k] if k < hp.n_base else hp.ext[k - hp.n_base]
        # what this rank uploads of EVERY column: all rows on one GPU, its block-cyclic pieces (+ the OOD reach) on several
        ranges = hp.prover.trace_rows_needed()
        # host side through the C ABI, as the Rust host would do it: the trace lives in ordinary host arrays that are pinned
        # once with ss_host_register (ministark's GpuVec columns) and uploaded with ss_memcpy_h2d on the copy stream
        import ctypes

        lib, h = ctx.lib, ctx.handle
        plan = []
        for k in range(n_trace_cols):
            parts = []
            for a, cnt in ranges:
                dst = dev_col(k)[a:a + cnt]
                host = dst.cpu().numpy()
                ctx.check(lib.ss_host_register(h, ctypes.c_void_p(host.ctypes.data), host.nbytes))
                parts.append((dst, host))
            plan.append(parts)
        my_h2d = sum(host.nbytes for col in plan for _, host in col)

        def e2e_step():
            # the trace is uploaded column by column on a copy stream; the LDE of column k waits only for column k,
            # so most of the H2D traffic overlaps the first LDE stage
            main = torch.cuda.current_stream()
            copy_stream.wait_stream(main)                  # the previous step has finished reading the buffers
            events = []
            for k in range(n_trace_cols):
                for dst, host in plan[k]:
                    ctx.check(lib.ss_memcpy_h2d(h, ctypes.c_void_p(dst.data_ptr()), ctypes.c_void_p(host.ctypes.data), host.nbytes,
                                                ctypes.c_void_p(copy_stream.cuda_stream)))
                ev = torch.cuda.Event()
                ev.record(copy_stream)
                events.append(ev)
            hp.column_ready = lambda k: main.wait_stream(copy_stream) if k is None else main.wait_event(events[k])
            hp.step()                                      # reads its roots, OOD values, remainder and openings back itself (D2H)
            main.wait_stream(copy_stream)

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
        hp.column_ready = None
        e2e_ms = e0.elapsed_time(e1) / args.steps
        t = torch.tensor([e2e_ms, float(my_h2d)], dtype=torch.float64, device="cuda")
        tm = t.clone()
        if world > 1:
            torch.distributed.all_reduce(tm, op=torch.distributed.ReduceOp.MAX)
            torch.distributed.all_reduce(t)
        e2e_ms, h2d = float(tm[0]), int(t[1])              # max over ranks; bytes summed over the ranks
        r = hp.last
        d2h = 32 * (3 + len(r.fri_roots)) + 32 * (len(r.ood_trace) + len(r.ood_composition)) + r.remainder.nbytes + r.opened_bytes
        e2e = {"value": hp.ntt_field_ops() / (e2e_ms * 1e-3), "unit": "field-ops/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms, "prove_seconds": e2e_ms / 1e3,
               "note": "the step's NTT field-ops over the WHOLE step incl. copies (every stage in the denominator).  The uploads run on a copy "
                       "stream under the first LDE stage, so this can land within run-to-run noise of the device-resident step."}
        for col in plan:
            for _, host in col:
                lib.ss_host_unregister(h, ctypes.c_void_p(host.ctypes.data))
        del plan

    # ---- the 2^24-point NTT the north_star roofline target is quoted on (8 columns, forward + inverse) ----------------
    prog = hp.prover._composition_program
    n_base, n_ext, N = hp.n_base, hp.n_ext, hp.N
    algo_ntt, compile_s = hp.ntt_algo_bytes(), hp.compile_s
    hp.base = hp.ext = hp.prover = hp.last = None
    del hp
    torch.cuda.empty_cache()
    nt_ops, nt_gbs, nt_sec = ntt_microbench(torch, ss, 24, 8, 3)

    if rank != 0:
        return
    peak, peak_src = peaks()
    achieved = algo_ntt / world / (ntt_ms * 1e-3) / 1e9
    # dominant kernel of the step: the composition-constraint evaluation (its stage is its launch)
    roofline = None
    if stages.get("constraint_eval"):
        ce_ms = stages["constraint_eval"]
        n_cols_read = n_base + n_ext + 1                                   # trace columns + the w = 1/(x-1) column
        rows = N // world
        algo = 32.0 * (n_cols_read + 1) * rows                              # every column element once + one output per row
        ce_ach = algo / (ce_ms * 1e-3) / 1e9
        ce_traffic = None
        cp = os.path.join(ROOT, "profiles", "ce_kernel_traffic.json")
        if os.path.exists(cp) and layout == "starknet":
            ce_traffic = json.load(open(cp))["dram_bytes_per_row"] * rows
        muls = prog.n_mul * rows / (ce_ms * 1e-3)
        roofline = {"bound": "hbm", "kernel": f"ce_gen_{layout}_composition (ss_constraint_eval)", "achieved": ce_ach, "peak": peak, "unit": "GB/s",
                    "frac": ce_ach / peak, "traffic": ce_traffic, "peak_source": peak_src, "launch_ms": ce_ms,
                    "algorithmic_bytes_per_row": 32 * (n_cols_read + 1),
                    "field_muls_per_s": muls, "field_mul_pipe_peak_per_s": MUL_PEAK, "field_mul_pipe_frac": muls / MUL_PEAK,
                    "field_mul_sustained_per_s": MUL_SUSTAINED, "field_mul_sustained_frac": muls / MUL_SUSTAINED,
                    "note": "arithmetic-bound: %d Montgomery multiplications + %d add/sub per row on %d algorithmic bytes; the integer pipe, not HBM, is the roof "
                            "(mul peak = 592 SMSPs x 1.965 GHz / 275 cycles per warp-multiplication, profiles/r01_ntt_tile_ab.md)" % (prog.n_mul, prog.n_addsub, 32 * (n_cols_read + 1))}
    cpu = None
    if not args.no_cpu:
        cpu = CpuArm(layout, tree, min(log_n, args.cpu_log_n)).run(1, 1)
    line = {
        "metric": "ntt_field_ops_per_s", "value": value, "unit": "field-ops/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "prove_seconds": ms_per_step / 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "u256 (Fp252 Montgomery, 8 x u32 limbs)", "data": "synthetic",
        "config": {"workload": workload_name(layout, log_n, tree, n_base, n_ext), "requested_log_steps": log_steps,
                   "parallelism": (f"{world} ranks: every column's transforms row-sharded (size-W transform across ranks + local size-n/W transforms, two NCCL "
                                   "all-to-alls per LDE column); hashing, constraint evaluation, DEEP on each rank's block-cyclic row pieces; sub-roots and partial "
                                   "OOD sums all-gathered; FRI on the gathered DEEP evaluations") if world > 1 else "1 rank",
                   "l2": "inputs_larger_than_L2", "stages_in_step": list(stages.keys()), "not_in_step": [],
                   "air": f"{layout} layout, {L.n_constraints} constraints (sandstorm_b200/air/layouts/{layout}.json)",
                   "template_compile_s_outside_step": round(compile_s, 1),
                   "patch_ms_inside_step": round(stages.get("patch", 0.0), 2)},
        "stages_ms": {k: round(v, 3) for k, v in stages.items()},
        "gpu_launches": launches,
        "clocks": clocks,
        "roofline": roofline,
        "roofline_ntt": {"bound": "hbm", "kernel": "ss::ntt_pass_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src, "basis": "SURVEY §8(d): n*s + N*s per LDE column, 2*N*s per NTT (not per pass)",
                         "ntt_2p24": {"field_ops_per_s": nt_ops, "achieved_gbs": nt_gbs, "frac": nt_gbs / peak, "ms_per_8_columns": nt_sec * 1e3,
                                      "note": "8 x 2^24-point transforms, forward + inverse averaged; the north_star target (>= 70 % of HBM) is quoted on this size"},
                         "note": "Fp252 NTT is bound by the carry-chained IMAD.WIDE pipe, not HBM (profiles/r01_pipe_microbench.md)"},
        "cpu_baseline": None if cpu is None else {"value": cpu["ntt_ops_per_s"], "unit": "field-ops/s", "cores": cpu["cores"], "kind": "port", "sample": cpu["sample"],
                                                  "prove_seconds_of_sample": cpu["prove_s"], "stages_s": cpu["stages_s"]},
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
    ap.add_argument("--workload", default="starknet22", choices=["starknet22", "recursive20", "ntt", "goldilocks"])
    ap.add_argument("--log-steps", type=int, default=0, help="log2 of Cairo steps (default: the workload's; n = 16 * steps)")
    ap.add_argument("--cpu-log-n", type=int, default=17, help="log2 rows of the bounded CPU sample (cpu_baseline / --impl reference)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        reference_arm(args)
    elif args.workload == "ntt":
        ntt_arm(args)
    elif args.workload == "goldilocks":
        goldilocks_arm(args)
    else:
        gpu_arm(args)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        if dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
