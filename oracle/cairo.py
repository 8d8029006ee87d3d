"""TEST INFRASTRUCTURE — never imported by the product (sandstorm_b200/) or by bench.py's timed path.

Pure-Python restatement of the reference's HOST side of a recursive-layout prove: the cairo-run file
parsers, the Cairo instruction word, the builtin instance traces, the execution-trace builder,
`build_extension_columns` and `gen_hints`.  It produces, from the reference's own fixture
(example/trace.bin, memory.bin, air-public-input.json), the exact matrix the reference hands to the hot
path, so that the transpiled AIR and the GPU stages can be checked on a trace that really satisfies the
constraints (random columns never do).  Follows, function by function:

    binary/src/lib.rs:148-162   RegisterStates::from_reader        -> read_register_states
    binary/src/lib.rs:173-213   Memory::from_reader                -> read_memory
    binary/src/lib.rs:310-340   AirPublicInput                     -> AirPublicInput
    binary/src/lib.rs:567-721   Word::{get_flag_prefix, get_*}     -> Word
    builtins/src/pedersen/mod.rs:70-181      InstanceTrace, gen_element_steps
    builtins/src/range_check/mod.rs:11-27    InstanceTrace
    builtins/src/bitwise/mod.rs:21-133       InstanceTrace, Partition*, dilute
    layouts/src/utils.rs:14-154, 226-393     public memory quotient, diluted cumulative value, ordered memory, pools
    layouts/src/recursive/trace.rs:88-689    ExecutionTrace::new    -> RecursiveTrace.__init__
    layouts/src/recursive/trace.rs:699-814   build_extension_columns
    layouts/src/recursive/air.rs:1202-1261   gen_hints

All values are canonical integers mod P (the reference holds Montgomery limbs; conversion happens where
the matrix is uploaded).  Columns are Python lists, column-major like ministark::Matrix."""
from __future__ import annotations

import json
import struct

P = 2**251 + 17 * 2**192 + 1

# layouts/src/recursive/mod.rs:16-35
CYCLE_HEIGHT = 16
PUBLIC_MEMORY_STEP = 16
MEMORY_STEP = 2
RANGE_CHECK_STEP = 4
PEDERSEN_BUILTIN_RATIO = 128
RANGE_CHECK_BUILTIN_RATIO = 8
RANGE_CHECK_BUILTIN_PARTS = 8
DILUTED_CHECK_N_BITS = 16
DILUTED_CHECK_SPACING = 4
BITWISE_RATIO = 8

# column enums of layouts/src/recursive/air.rs:1266-1729 (value = row shift inside the enum's step)
NPC = dict(Pc=0, Instruction=1, PubMemAddr=2, PubMemVal=3, MemOp0Addr=4, MemOp0=5, PedersenInput0Addr=10, PedersenInput0Val=11,
           PedersenInput1Addr=1034, PedersenInput1Val=1035, PedersenOutputAddr=522, PedersenOutputVal=523, RangeCheck128Addr=74,
           RangeCheck128Val=75, BitwisePoolAddr=26, BitwisePoolVal=27, BitwiseXOrYAddr=42, BitwiseXOrYVal=43, MemDstAddr=8, MemDst=9,
           MemOp1Addr=12, MemOp1=13, UnusedAddr=14, UnusedVal=15)
RANGE_CHECK = dict(OffDst=0, Ordered=2, OffOp1=4, OffOp0=8, Unused=12)
AUXILIARY = dict(Ap=1, Tmp0=3, Op0MulOp1=5, Fp=9, Tmp1=11, Res=13)
PEDERSEN = dict(PartialSumX=1, PartialSumY=3, Suffix=0, Slope=2, Bit251AndBit196AndBit192=7, Bit251AndBit196=1022)
RC_BUILTIN_COMPONENT = 12
BITWISE_RES_SHIFTED = (1, 65, 33, 97)              # Bits16Chunk3Offset{0,1,2,3}ResShifted
# challenge indices, air.rs:1759-1806
MEM_Z, MEM_A, RC_Z, DILUTED_PERM_Z, DILUTED_AGG_Z, DILUTED_AGG_A = 0, 1, 2, 3, 4, 5
# Flag enum, binary/src/lib.rs:737-772
FLAGS = dict(DstReg=0, Op0Reg=1, Op1Imm=2, Op1Fp=3, Op1Ap=4, ResAdd=5, ResMul=6, PcJumpAbs=7, PcJumpRel=8, PcJnz=9, ApAdd=10, ApAdd1=11,
             OpcodeCall=12, OpcodeRet=13, OpcodeAssertEq=14, Zero=15)
HALF_OFFSET = 2**15


# ---- binary/src/lib.rs ------------------------------------------------------------------------------------
def read_register_states(path):
    """bincode of RegisterState {ap, fp, pc} (usize = u64 LE each), lib.rs:52-56,152-161.  Returns [(ap, fp, pc)]."""
    raw = open(path, "rb").read()
    assert len(raw) % 24 == 0
    return [struct.unpack_from("<QQQ", raw, o) for o in range(0, len(raw), 24)]


def read_memory(path):
    """u64 LE address + 32-byte LE word per entry, lib.rs:177-212.  Returns a list indexed by address (None = hole)."""
    raw = open(path, "rb").read()
    assert len(raw) % 40 == 0
    entries = [(struct.unpack_from("<Q", raw, o)[0], int.from_bytes(raw[o + 8:o + 40], "little")) for o in range(0, len(raw), 40)]
    mem = [None] * (max(a for a, _ in entries) + 1)
    for a, w in entries:
        assert w < P
        mem[a] = w
    return mem


class AirPublicInput:
    """lib.rs:310-340."""

    def __init__(self, d: dict):
        self.rc_min, self.rc_max, self.n_steps, self.layout = d["rc_min"], d["rc_max"], d["n_steps"], d["layout"]
        self.memory_segments = {k: (v["begin_addr"], v["stop_ptr"]) for k, v in d["memory_segments"].items()}
        self.public_memory = [(e["address"], int(e["value"], 16)) for e in d["public_memory"]]

    @classmethod
    def from_file(cls, path):
        return cls(json.load(open(path)))

    def initial_pc(self): return self.memory_segments["program"][0]
    def final_pc(self): return self.memory_segments["program"][1]
    def initial_ap(self): return self.memory_segments["execution"][0]
    def final_ap(self): return self.memory_segments["execution"][1]

    def public_memory_padding(self):
        return next(e for e in self.public_memory if e[0] == 1)


class Word:
    """A Cairo instruction word (lib.rs:567-721)."""

    def __init__(self, w: int):
        self.w = w

    def get_flag(self, f): return (self.w >> (48 + f)) & 1
    def get_flag_prefix(self, f): return 0 if f == 15 else (self.w >> (48 + f)) & ((1 << (15 - f)) - 1)
    def get_off_dst(self): return self.w & 0xFFFF
    def get_off_op0(self): return (self.w >> 16) & 0xFFFF
    def get_off_op1(self): return (self.w >> 32) & 0xFFFF
    def get_op0_addr(self, ap, fp): return self.get_off_op0() + (fp if self.get_flag(FLAGS["Op0Reg"]) else ap) - HALF_OFFSET
    def get_dst_addr(self, ap, fp): return self.get_off_dst() + (fp if self.get_flag(FLAGS["DstReg"]) else ap) - HALF_OFFSET

    def flag_group(self, name):
        g = self.get_flag
        F = FLAGS
        return {"Op1Src": g(F["Op1Imm"]) + 2 * g(F["Op1Fp"]) + 4 * g(F["Op1Ap"]), "ResLogic": g(F["ResAdd"]) + 2 * g(F["ResMul"]),
                "PcUpdate": g(F["PcJumpAbs"]) + 2 * g(F["PcJumpRel"]) + 4 * g(F["PcJnz"]), "ApUpdate": g(F["ApAdd"]) + 2 * g(F["ApAdd1"]),
                "Opcode": g(F["OpcodeCall"]) + 2 * g(F["OpcodeRet"]) + 4 * g(F["OpcodeAssertEq"])}[name]

    def get_op1_addr(self, pc, ap, fp, mem):
        src = self.flag_group("Op1Src")
        base = {0: lambda: mem[self.get_op0_addr(ap, fp)], 1: lambda: pc, 2: lambda: fp, 4: lambda: ap}[src]()
        return self.get_off_op1() + base - HALF_OFFSET

    def get_op0(self, ap, fp, mem): return mem[self.get_op0_addr(ap, fp)]
    def get_dst(self, ap, fp, mem): return mem[self.get_dst_addr(ap, fp)]
    def get_op1(self, pc, ap, fp, mem): return mem[self.get_op1_addr(pc, ap, fp, mem)]

    def get_res(self, pc, ap, fp, mem):
        pc_update, res_logic = self.flag_group("PcUpdate"), self.flag_group("ResLogic")
        if pc_update == 4:
            assert res_logic == 0 and self.flag_group("Opcode") == 0 and self.flag_group("ApUpdate") != 1
            d = self.get_dst(ap, fp, mem)
            return pow(d, -1, P) if d else 0                    # res holds dst^-1 for jnz (lib.rs:668-676)
        assert pc_update in (0, 1, 2)
        op0, op1 = self.get_op0(ap, fp, mem), self.get_op1(pc, ap, fp, mem)
        return {0: op1, 1: (op0 + op1) % P, 2: op0 * op1 % P}[res_logic]

    def get_tmp0(self, ap, fp, mem): return self.get_dst(ap, fp, mem) if self.get_flag(FLAGS["PcJnz"]) else 0
    def get_tmp1(self, pc, ap, fp, mem): return self.get_tmp0(ap, fp, mem) * self.get_res(pc, ap, fp, mem) % P


# ---- StarkWare curve / Pedersen builtin (builtins/src/utils.rs:122-183, pedersen/mod.rs, constants.rs:6-29) ---------
PEDERSEN_POINTS = [
    (2089986280348253421170679821480865132823066470938446095505822317253594081284, 1713931329540660377023406109199410414810705867260802078187082345529207694986),
    (996781205833008774514500082376783249102396023663454813447423147977397232763, 1668503676786377725805489344771023921079126552019160156920634619255970485781),
    (2251563274489750535117886426533222435294046428347329203627021249169616184184, 1798716007562728905295480679789526322175868328062420237419143593021674992973),
    (2138414695194151160943305727036575959195309218611738193261179310511854807447, 113410276730064486255102093846540133784865286929052426931474106396135072156),
    (2379962749567351885752724891227938183011949129833673362440656643086021394946, 776496453633298175483985398648758586525933812536653089401905292063708816422),
]


def ec_slope(p1, p2):
    """calculate_slope (utils.rs:159-182); curve y^2 = x^3 + x + b."""
    (x1, y1), (x2, y2) = p1, p2
    if x1 == x2:
        assert y1 == y2
        return (3 * x1 * x1 + 1) * pow(2 * y1, -1, P) % P
    return (y2 - y1) * pow(x2 - x1, -1, P) % P


def ec_add(p1, p2):
    lam = ec_slope(p1, p2)
    x3 = (lam * lam - p1[0] - p2[0]) % P
    return x3, (lam * (p1[0] - x3) - p1[1]) % P


_doubling_chains: dict = {}


def _constant_points(p1, p2):
    key = (p1, p2)
    if key not in _doubling_chains:
        pts, acc = [], p1
        for _ in range(248):
            pts.append(acc)
            acc = ec_add(acc, acc)
        acc = p2
        for _ in range(4):
            pts.append(acc)
            acc = ec_add(acc, acc)
        _doubling_chains[key] = pts
    return _doubling_chains[key]


def pedersen_element_steps(x, p0, p1, p2):
    """gen_element_steps (pedersen/mod.rs:133-181): 256 steps of (partial point, suffix, slope)."""
    pts, partial, out = _constant_points(p1, p2), p0, []
    for i in range(256):
        suffix = x >> i
        slope, nxt = 0, partial
        if suffix & 1:
            slope = ec_slope(pts[i], partial)
            nxt = ec_add(partial, pts[i])
        out.append((partial, suffix, slope))
        partial = nxt
    return out, partial


_pedersen_traces: dict = {}


def pedersen_instance_trace(a, b):
    """InstanceTrace::new (pedersen/mod.rs:83-130).  Returns dict(steps[512], output, a/b bit flags)."""
    if (a, b) not in _pedersen_traces:
        P0, P1, P2, P3, P4 = PEDERSEN_POINTS
        a_steps, after_a = pedersen_element_steps(a, P0, P1, P2)
        assert a_steps[-1][0] == after_a                     # the top bits of a field element are zero
        b_steps, after_b = pedersen_element_steps(b, after_a, P3, P4)
        bit = lambda v, k: (v >> k) & 1
        _pedersen_traces[(a, b)] = dict(
            steps=a_steps + b_steps, output=b_steps[-1][0][0],
            a_flags=(bit(a, 251) & bit(a, 196) & bit(a, 192), bit(a, 251) & bit(a, 196)),
            b_flags=(bit(b, 251) & bit(b, 196) & bit(b, 192), bit(b, 251) & bit(b, 196)))
    return _pedersen_traces[(a, b)]


def pedersen_hash(a, b):
    return pedersen_instance_trace(a, b)["output"]


# ---- bitwise builtin (builtins/src/bitwise/mod.rs) -------------------------------------------------------
def partition64(v):
    """Partition64<4>::new: segment s keeps bits b*4+s of v, at position b*4."""
    seg = [0, 0, 0, 0]
    for bb in range(16):
        for s in range(4):
            seg[s] |= ((v >> (bb * 4 + s)) & 1) << (bb * 4)
    return seg


def partition256(v):
    """Partition256::new: [low.low, low.high, high.low, high.high], each 4 segments."""
    return [partition64((v >> (64 * k)) & (2**64 - 1)) for k in range(4)]


def dilute(v):
    r = 0
    for i in range(64):
        r |= ((v >> i) & 1) << (i * DILUTED_CHECK_SPACING)
    return r


def undilute(v):
    """DilutedCheckPool::push_diluted (layouts/src/utils.rs:255-272)."""
    assert v & ~sum(1 << (i * DILUTED_CHECK_SPACING) for i in range(DILUTED_CHECK_N_BITS)) == 0
    r = 0
    for i in range(DILUTED_CHECK_N_BITS):
        r |= ((v >> (i * DILUTED_CHECK_SPACING)) & 1) << i
    return r


# ---- pools (layouts/src/utils.rs:226-393) -----------------------------------------------------------------
def ordered_with_padding(vals, lo=None, hi=None):
    """{RangeCheck,DilutedCheck}Pool::get_ordered_values_with_padding -> (ordered incl. padding, padding)."""
    if not vals:
        return [], list(range(lo, hi + 1))
    s = sorted(vals)
    pad = []
    if lo is not None:
        assert s[0] >= lo and s[-1] <= hi
        pad += list(range(lo, s[0])) + list(range(s[-1] + 1, hi + 1))
    for a, b in zip(s, s[1:]):
        pad.extend(range(a + 1, b))
    return sorted(s + pad), pad


def get_ordered_memory_accesses(trace_len, accesses, public_memory, padding):
    """layouts/src/utils.rs:116-154."""
    cells = trace_len // PUBLIC_MEMORY_STEP
    ordered = sorted(accesses + [padding] * (cells - len(public_memory)) + public_memory, key=lambda e: e[0])
    zeros, ordered = ordered[:cells], ordered[cells:]
    assert all(a == 0 for a, _ in zeros) and ordered[0][0] == 1
    for cur, nxt in zip(ordered, ordered[1:]):
        assert cur == nxt or cur[0] == nxt[0] - 1, (cur, nxt)
    return ordered


def compute_public_memory_quotient(z, alpha, trace_len, public_memory, padding):
    """layouts/src/utils.rs:14-46."""
    s, k = trace_len // PUBLIC_MEMORY_STEP, len(public_memory)
    den = 1
    for a, v in public_memory:
        den = den * (z - (alpha * v + a)) % P
    pad = pow((z - (alpha * padding[1] + padding[0])) % P, s - k, P)
    return pow(z, s, P) * pow(den * pad % P, -1, P) % P


def compute_diluted_cumulative_value(z, alpha, n_bits=DILUTED_CHECK_N_BITS, spacing=DILUTED_CHECK_SPACING):
    """layouts/src/utils.rs:83-110."""
    diff_mult, diff_x = 1 << spacing, (1 << spacing) - 2
    p, q, x = (z + 1) % P, 1, 1
    for _ in range(1, n_bits):
        x = (x + diff_x) % P
        diff_x = diff_x * diff_mult % P
        xp = x * p % P
        y = (p + z * xp) % P
        q = (q + q * y + x * xp) % P
        p = p * y % P
    return (p + q * alpha) % P


def batch_inverse(vals):
    """ark_ff::batch_inversion (Montgomery's trick); every value is non-zero here."""
    pref, acc = [], 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % P
    inv = pow(acc, -1, P)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % P
        inv = inv * vals[i] % P
    return out


# ---- layouts/src/recursive/trace.rs ------------------------------------------------------------------------
class RecursiveTrace:
    """ExecutionTrace::new (trace.rs:88-689).  private_input: dict with lists "pedersen" [(index, a, b)], "range_check"
    [(index, value)], "bitwise" [(index, x, y)] (binary/src/lib.rs:521-535); the committed example has all three empty."""

    def __init__(self, public_input: AirPublicInput, register_states, memory, private_input=None):
        priv = private_input or {"pedersen": [], "range_check": [], "bitwise": []}
        self.public_input = public_input
        num_cycles = len(register_states)
        assert num_cycles & (num_cycles - 1) == 0
        n = self.trace_len = num_cycles * CYCLE_HEIGHT
        self.public_memory = list(public_input.public_memory)
        padding = self.padding_entry = public_input.public_memory_padding()
        words = {}

        def word_at(pc):
            if pc not in words:
                words[pc] = Word(memory[pc])
            return words[pc]

        flags = [0] * n
        npc = [padding[0], padding[1]] * (n // 2)                 # every memory cell defaults to the padding entry (:121-131)
        # 16-bit range-check pool: instruction offsets, then the parts of the 128-bit range-check builtin (:133-153)
        rc_pool = []
        for ap, fp, pc in register_states:
            w = word_at(pc)
            rc_pool += [w.get_off_dst(), w.get_off_op0(), w.get_off_op1()]
        rc128 = [(idx, val, self._rc_parts(val)) for idx, val in priv["range_check"]]
        for _, _, parts in rc128:
            rc_pool += parts
        ordered_rc, rc_padding = ordered_with_padding(rc_pool)
        self.range_check_min, self.range_check_max = min(rc_pool), max(rc_pool)
        rc_max = self.range_check_max
        ordered_rc, rc_padding = iter(ordered_rc), iter(rc_padding)
        rc_col = [rc_max] * n
        aux = [0] * n
        for c, (ap, fp, pc) in enumerate(register_states):             # :172-236
            w, o = word_at(pc), c * CYCLE_HEIGHT
            assert not w.get_flag(FLAGS["Zero"])
            op0, op1, dst = w.get_op0(ap, fp, memory), w.get_op1(pc, ap, fp, memory), w.get_dst(ap, fp, memory)
            for f in range(16):
                flags[o + f] = w.get_flag_prefix(f)
            npc[o + NPC["Pc"]], npc[o + NPC["Instruction"]] = pc, w.w
            npc[o + NPC["MemOp0Addr"]], npc[o + NPC["MemOp0"]] = w.get_op0_addr(ap, fp), op0
            npc[o + NPC["MemDstAddr"]], npc[o + NPC["MemDst"]] = w.get_dst_addr(ap, fp), dst
            npc[o + NPC["MemOp1Addr"]], npc[o + NPC["MemOp1"]] = w.get_op1_addr(pc, ap, fp, memory), op1
            npc[o + NPC["PubMemAddr"]] = npc[o + NPC["PubMemVal"]] = 0
            rc_col[o + RANGE_CHECK["OffDst"]], rc_col[o + RANGE_CHECK["OffOp1"]], rc_col[o + RANGE_CHECK["OffOp0"]] = \
                w.get_off_dst(), w.get_off_op1(), w.get_off_op0()
            aux[o + AUXILIARY["Tmp0"]], aux[o + AUXILIARY["Tmp1"]] = w.get_tmp0(ap, fp, memory), w.get_tmp1(pc, ap, fp, memory)
            aux[o + AUXILIARY["Ap"]], aux[o + AUXILIARY["Fp"]] = ap, fp
            aux[o + AUXILIARY["Op0MulOp1"]], aux[o + AUXILIARY["Res"]] = op0 * op1 % P, w.get_res(pc, ap, fp, memory)
        # dummy 128-bit range checks built from the 16-bit padding values (:238-255)
        for index in range(len(rc128), num_cycles // RANGE_CHECK_BUILTIN_RATIO):
            value = 0
            for _ in range(RANGE_CHECK_BUILTIN_PARTS):
                value = (value << 16) + next(rc_padding, rc_max)
            rc128.append((index, value, self._rc_parts(value)))
        for cycle in range(num_cycles):                                # :257-283
            o = cycle * CYCLE_HEIGHT
            if cycle % 2 == 1:
                rc_col[o + RANGE_CHECK["Unused"]] = next(rc_padding, rc_max)
            for off in range(0, CYCLE_HEIGHT, RANGE_CHECK_STEP):
                rc_col[o + off + RANGE_CHECK["Ordered"]] = next(ordered_rc, rc_max)
        assert next(rc_padding, None) is None and next(ordered_rc, None) is None
        diluted_ordered, diluted_unordered = [0] * n, [0] * n
        # Pedersen builtin: one hash per 2048 rows (:305-372)
        step_rows = PEDERSEN_BUILTIN_RATIO * CYCLE_HEIGHT
        part_rows = step_rows // 512
        ped_begin = self.initial_pedersen_address = public_input.memory_segments["pedersen"][0]
        instances = list(priv["pedersen"])
        for k in range(n // step_rows):
            index, a, b = instances[k] if k < len(instances) else (k, 0, 0)
            t, base = pedersen_instance_trace(a % P, b % P), k * step_rows
            for s, (pt, suffix, slope) in enumerate(t["steps"]):
                r = base + s * part_rows
                rc_col[r + PEDERSEN["PartialSumX"]], rc_col[r + PEDERSEN["PartialSumY"]] = pt
                aux[r + PEDERSEN["Suffix"]], aux[r + PEDERSEN["Slope"]] = suffix % P, slope
            for half, (f3, f2) in ((0, t["a_flags"]), (1, t["b_flags"])):
                hb = base + half * part_rows * 256
                aux[hb + PEDERSEN["Bit251AndBit196"]] = f2
                aux[hb + PEDERSEN["Bit251AndBit196AndBit192"]] = f3
            addr = ped_begin + index * 3
            npc[base + NPC["PedersenInput0Addr"]], npc[base + NPC["PedersenInput0Val"]] = addr, a % P
            npc[base + NPC["PedersenInput1Addr"]], npc[base + NPC["PedersenInput1Val"]] = addr + 1, b % P
            npc[base + NPC["PedersenOutputAddr"]], npc[base + NPC["PedersenOutputVal"]] = addr + 2, t["output"]
        # range-check builtin: one 128-bit value per 128 rows (:374-411)
        rc_rows = RANGE_CHECK_BUILTIN_RATIO * CYCLE_HEIGHT
        rc_part_rows = rc_rows // RANGE_CHECK_BUILTIN_PARTS
        rc_begin = self.initial_rc_address = public_input.memory_segments["range_check"][0]
        for k in range(n // rc_rows):
            index, value, parts = rc128[k]
            base = k * rc_rows
            for j, part in enumerate(parts):
                rc_col[base + RC_BUILTIN_COMPONENT + rc_part_rows * j] = part
            npc[base + NPC["RangeCheck128Addr"]], npc[base + NPC["RangeCheck128Val"]] = rc_begin + index, value % P
        # bitwise builtin: one instance per 128 rows, diluted chunks into the diluted-check column (:413-540)
        bw_rows = BITWISE_RATIO * CYCLE_HEIGHT
        bw_begin = self.initial_bitwise_address = public_input.memory_segments["bitwise"][0]
        bw_instances, diluted_pool = list(priv["bitwise"]), []
        dummy = None
        for k in range(n // bw_rows):
            index, x, y = bw_instances[k] if k < len(bw_instances) else (k, 0, 0)
            base = k * bw_rows
            if (x, y) == (0, 0) and dummy is not None:                 # every dummy instance contributes the same zeros
                diluted_pool += dummy
            else:
                before = len(diluted_pool)
                parts = [partition256(v) for v in (x, y, x & y, x ^ y)]
                v = [parts[2][3][j] + parts[3][3][j] for j in range(4)]     # x&y + x^y of the top 64-bit chunk (:437-462)
                for j, sh in enumerate((4, 4, 4, 8)):
                    assert v[j] == ((v[j] << sh) & (2**64 - 1)) >> sh
                    s = v[j] << sh
                    diluted_pool.append(undilute(s))
                    diluted_unordered[base + BITWISE_RES_SHIFTED[j]] = s
                for q, part in enumerate(parts):                           # order matters: x, y, x&y, x^y (:464-501)
                    for chunk in range(4):
                        for j in range(4):
                            diluted_unordered[base + 32 * q + 8 * chunk + 2 * j] = part[chunk][j]
                            diluted_pool.append(undilute(part[chunk][j]))
                if (x, y) == (0, 0):
                    dummy = diluted_pool[before:]
            addr_step = bw_rows // 4
            o = base + NPC["BitwisePoolAddr"]
            for j, val in enumerate((x, y, x & y, x ^ y)):
                npc[o + addr_step * j], npc[o + addr_step * j + 1] = bw_begin + index * 5 + j, val % P
            npc[base + NPC["BitwiseXOrYAddr"]], npc[base + NPC["BitwiseXOrYVal"]] = bw_begin + index * 5 + 4, (x | y) % P
        ordered_dil, dil_padding = ordered_with_padding(diluted_pool, 0, (1 << DILUTED_CHECK_N_BITS) - 1)
        ordered_dil, dil_padding = [dilute(v) for v in ordered_dil], iter(dilute(v) for v in dil_padding)
        done = False                                                       # padding into the free odd cells (:563-587)
        for base in range(0, n, bw_rows):
            for off in range(1, bw_rows, 2):
                if off in BITWISE_RES_SHIFTED:
                    continue
                v = next(dil_padding, None)
                if v is None:
                    done = True
                    break
                diluted_unordered[base + off] = v
            if done:
                break
        diluted_ordered[n - len(ordered_dil):] = ordered_dil                   # :589-593
        assert next(dil_padding, None) is None
        # memory: fill address gaps through the unused cells, then sort (:599-649)
        accesses = [(npc[i], npc[i + 1]) for i in range(0, n, 2)]
        srt = sorted(accesses + self.public_memory, key=lambda e: e[0])
        gaps = []
        for (a, _), (b, _) in zip(srt, srt[1:]):
            gaps.extend(range(a + 1, b))
        gaps = iter(gaps)
        for o in range(0, n, CYCLE_HEIGHT):
            addr = next(gaps, None)
            if addr is None:
                break
            npc[o + NPC["UnusedAddr"]], npc[o + NPC["UnusedVal"]] = addr, 0
        assert next(gaps, None) is None
        accesses = [(npc[i], npc[i + 1]) for i in range(0, n, 2)]
        ordered_mem = get_ordered_memory_accesses(n, accesses, self.public_memory, padding)
        memory_col = [v for e in ordered_mem for v in e]
        assert len(memory_col) == n
        # base matrix, column order of trace.rs:652-660
        self.base_columns = [flags, diluted_unordered, diluted_ordered, npc, memory_col, rc_col, aux]
        self.initial_registers, self.final_registers = register_states[0], register_states[-1]

    @staticmethod
    def _rc_parts(value):
        """range_check::InstanceTrace::new: most significant 16-bit part first."""
        assert value < 1 << (16 * RANGE_CHECK_BUILTIN_PARTS)
        return [(value >> ((RANGE_CHECK_BUILTIN_PARTS - i - 1) * 16)) & 0xFFFF for i in range(RANGE_CHECK_BUILTIN_PARTS)]

    def build_extension_columns(self, challenges):
        """trace.rs:699-814.  Returns [diluted aggregate, diluted permutation, memory + range-check permutation]."""
        n = self.trace_len
        _, dil_unordered, dil_ordered, npc, mem, rc, _ = self.base_columns

        def running_quotient(num_factors, den_factors):
            nums, dens, a, b = [], [], 1, 1
            for f, g in zip(num_factors, den_factors):
                a, b = a * f % P, b * g % P
                nums.append(a)
                dens.append(b)
            return [x * y % P for x, y in zip(nums, batch_inverse(dens))], a, b

        z, alpha = challenges[MEM_Z], challenges[MEM_A]
        mem_perm, _, _ = running_quotient(((z - (alpha * npc[i + 1] + npc[i])) % P for i in range(0, n, MEMORY_STEP)),
                                          ((z - (alpha * mem[i + 1] + mem[i])) % P for i in range(0, n, MEMORY_STEP)))
        z = challenges[RC_Z]
        rc_perm, a, b = running_quotient(((z - rc[i + RANGE_CHECK["OffDst"]]) % P for i in range(0, n, RANGE_CHECK_STEP)),
                                         ((z - rc[i + RANGE_CHECK["Ordered"]]) % P for i in range(0, n, RANGE_CHECK_STEP)))
        assert a == b
        z = challenges[DILUTED_PERM_Z]
        dil_perm, a, b = running_quotient(((z - v) % P for v in dil_unordered), ((z - v) % P for v in dil_ordered))
        assert a == b
        perm = [0] * n
        perm[0::MEMORY_STEP] = mem_perm                               # Permutation::Memory = (col 9, shift 0)
        perm[1::RANGE_CHECK_STEP] = rc_perm                           # Permutation::RangeCheck = (col 9, shift 1)
        z, alpha = challenges[DILUTED_AGG_Z], challenges[DILUTED_AGG_A]
        agg, acc = [1] * n, 1
        for i in range(1, n):
            u = (dil_ordered[i] - dil_ordered[i - 1]) % P
            acc = (acc * (1 + z * u) + alpha * u * u) % P
            agg[i] = acc
        return [agg, dil_perm, perm]

    def gen_hints(self, challenges):
        """AirConfig::gen_hints (recursive/air.rs:1202-1261), hint order = PublicInputHint (:1732-1747)."""
        pi = self.public_input
        quotient = compute_public_memory_quotient(challenges[MEM_Z], challenges[MEM_A], self.trace_len, self.public_memory, self.padding_entry)
        cumulative = compute_diluted_cumulative_value(challenges[DILUTED_AGG_Z], challenges[DILUTED_AGG_A])
        return [pi.initial_ap(), pi.initial_pc(), pi.final_ap(), pi.final_pc(), quotient, 1, pi.rc_min, pi.rc_max, 1, 0, cumulative,
                pi.memory_segments["pedersen"][0], pi.memory_segments["range_check"][0], pi.memory_segments["bitwise"][0]]


def load_example(directory):
    """The reference's example/ fixture (array-sum, recursive layout, 16384 steps)."""
    import os

    pub = AirPublicInput.from_file(os.path.join(directory, "air-public-input.json"))
    regs = read_register_states(os.path.join(directory, "trace.bin"))
    mem = read_memory(os.path.join(directory, "memory.bin"))
    return RecursiveTrace(pub, regs, mem)


# =====================================================================================================================
# starknet layout (layouts/src/starknet/{mod,trace,air}.rs): the same CPU / memory / range-check / Pedersen / bitwise logic on
# the starknet cell map, plus the ECDSA, EC-op and Poseidon builtins.
EC_ORDER = 3618502788666131213697322783095070105526743751716087489154079457884512865583          # builtins/src/utils.rs:134
EC_GEN = (874739451078007766457464989774322083649278607533249481151382481072868806602, 152666792071518830868575557812948353041420400780739481342941381225525861407)
EC_BETA = 3141592653589793238462643383279502884197169399375105820974944592307816406665
SHIFT_POINT = PEDERSEN_POINTS[0]                                                                  # ecdsa/mod.rs SHIFT_POINT = P0

SN = dict(CYCLE=16, PUB_MEM_STEP=8, PEDERSEN_RATIO=32, RC_RATIO=16, ECDSA_RATIO=2048, BITWISE_RATIO=64, EC_OP_RATIO=1024, POSEIDON_RATIO=32,
          DILUTED_STEP=8)                                                                         # starknet/mod.rs:14-49
SN_NPC = dict(Pc=0, Instruction=1, PubMemAddr=2, PubMemVal=3, MemOp0Addr=4, MemOp0=5, PedersenInput0Addr=6, PedersenInput1Addr=262,
              PedersenOutputAddr=134, RangeCheck128Addr=70, EcdsaPubkeyAddr=390, EcdsaMessageAddr=16774, BitwisePoolAddr=198, BitwiseXOrYAddr=902,
              EcOpPXAddr=8582, EcOpPYAddr=4486, EcOpQXAddr=12678, EcOpQYAddr=2438, EcOpMAddr=10630, EcOpRXAddr=6534, EcOpRYAddr=14726,
              PoseidonInput0Addr=38, PoseidonInput1Addr=102, PoseidonInput2Addr=166, PoseidonOutput0Addr=230, PoseidonOutput1Addr=294,
              PoseidonOutput2Addr=358, MemDstAddr=8, MemDst=9, MemOp1Addr=12, MemOp1=13, UnusedAddr=14, UnusedVal=15)     # air.rs:2914-3101
SN_AUX = dict(Ap=0, Tmp0=2, Op0MulOp1=4, Fp=8, Tmp1=10, Res=12)
SN_ECDSA = dict(PubkeyDoublingX=1, PubkeyDoublingY=33, PubkeyDoublingSlope=35, PubkeyPartialSumX=17, PubkeyPartialSumY=49, PubkeyPartialSumXDiffInv=51,
                PubkeyPartialSumSlope=19, RSuffix=9, MessageSuffix=59, GeneratorPartialSumY=91, GeneratorPartialSumX=27, GeneratorPartialSumXDiffInv=7,
                GeneratorPartialSumSlope=123, RPointSlope=16331, RPointXDiffInv=32715, RInv=16355, WInv=32739, MessageInv=16363, PubkeyXSquared=32747,
                BSlope=32763, BXDiffInv=32647)                                                    # air.rs:2691-2782
SN_ECOP = dict(QDoublingX=41, QDoublingY=25, QDoublingSlope=57, RPartialSumX=5, RPartialSumY=37, RPartialSumSlope=11, RPartialSumXDiffInv=43, MSuffix=21,
               MBit251AndBit196AndBit192=16371, MBit251AndBit196=16339)
SN_POSEIDON = dict(Full0=53, Full0Sq=29, Full1=13, Full1Sq=61, Full2=45, Full2Sq=3, Partial0=3, Partial0Sq=7, Partial1=6, Partial1Sq=14)   # air.rs:2593-2602
SN_BITWISE_SHIFTED = (9, 521, 265, 777)


def ec_neg(p):
    return p[0], -p[1] % P


def ec_mul(k, pt):
    acc = None
    while k:
        if k & 1:
            acc = pt if acc is None else ec_add(acc, pt)
        pt = ec_add(pt, pt)
        k >>= 1
    return acc


def doubling_steps(num, pt):
    """ecdsa/mod.rs doubling_steps: (point, tangent slope) for num successive doublings."""
    out = []
    for _ in range(num):
        out.append((pt, ec_slope(pt, pt)))
        pt = ec_add(pt, pt)
    return out


def ec_mad_steps(x, point, shift, max_doublings):
    """gen_ec_mad_steps (ecdsa/mod.rs:166-211; ec_op/mod.rs:102-135 with max_doublings = 256): 256 steps of
    (partial sum, fixed point, suffix, slope, 1 / (partial.x - point.x)) for shift + x * point."""
    partial, out = shift, []
    for i in range(256):
        suffix = x >> i
        slope, nxt = 0, partial
        if suffix & 1:
            slope = ec_slope(point, partial)
            nxt = ec_add(partial, point)
        out.append((partial, point, suffix, slope, pow((partial[0] - point[0]) % P, -1, P)))
        partial = nxt
        if i < max_doublings:
            point = ec_add(point, point)
    return out


def mimic_ec_mad(m, point, shift):
    """mimic_ec_mad_air: shift + m * point the way the AIR computes it."""
    partial = shift
    while m:
        assert partial[0] != point[0]
        if m & 1:
            partial = ec_add(partial, point)
        point = ec_add(point, point)
        m >>= 1
    return partial


_ecdsa_dummy = None


def ecdsa_dummy_trace():
    """ecdsa::InstanceTrace::new(gen_dummy_instance(0)) (ecdsa/mod.rs:67-140, 229-275): private key 1, message pedersen(1, 0)."""
    global _ecdsa_dummy
    if _ecdsa_dummy is not None:
        return _ecdsa_dummy
    message = pedersen_hash(1, 0)
    assert 0 < message < 2**251
    k = 1
    while True:
        r = ec_mul(k, EC_GEN)[0]
        if 0 < r < 2**251 and (message + r) % EC_ORDER:
            w = k * pow((message + r) % EC_ORDER, -1, EC_ORDER) % EC_ORDER
            if 0 < w < 2**251:
                break
        k += 1
    pubkey, shift = EC_GEN, SHIFT_POINT
    zg = mimic_ec_mad(message, EC_GEN, ec_neg(shift))
    qr = mimic_ec_mad(r, pubkey, shift)
    b = ec_add(zg, qr)
    wb = mimic_ec_mad(w, b, shift)
    assert ec_add(wb, ec_neg(shift))[0] == r                          # the signature verifies
    t = dict(message=message, pubkey=pubkey, r=r, w=w,
             zg_steps=ec_mad_steps(message, EC_GEN, ec_neg(shift), 250), rq_steps=ec_mad_steps(r, pubkey, shift, 255), wb_steps=ec_mad_steps(w, b, shift, 255),
             pubkey_doubling=doubling_steps(256, pubkey), b_doubling=doubling_steps(256, b),
             b_slope=ec_slope(zg, qr), b_x_diff_inv=pow((zg[0] - qr[0]) % P, -1, P), w_inv=pow(w, -1, P), r_inv=pow(r, -1, P),
             message_inv=pow(message, -1, P), r_point_slope=ec_slope(wb, ec_neg(shift)), r_point_x_diff_inv=pow((wb[0] - shift[0]) % P, -1, P))
    assert t["zg_steps"][-1][0] == zg and t["rq_steps"][-1][0] == qr and t["wb_steps"][-1][0] == wb
    _ecdsa_dummy = t
    return t


_ec_op_dummy = None


def ec_op_dummy_trace():
    """ec_op::InstanceTrace::new(gen_dummy_instance(0)) (ec_op/mod.rs:35-100): r = P0 + 1 * G."""
    global _ec_op_dummy
    if _ec_op_dummy is None:
        p, q, m = PEDERSEN_POINTS[0], EC_GEN, 1
        steps = ec_mad_steps(m, q, p, 256)
        _ec_op_dummy = dict(p=p, q=q, m=m, r=mimic_ec_mad(m, q, p), r_steps=steps, q_doubling=doubling_steps(256, q))
        assert steps[-1][0] == _ec_op_dummy["r"]
    return _ec_op_dummy


_poseidon_cache: dict = {}


def poseidon_trace(inputs, params):
    """poseidon::InstanceTrace::new (poseidon/mod.rs:66-128)."""
    key = tuple(inputs)
    if key in _poseidon_cache:
        return _poseidon_cache[key]
    mds = lambda s: [(3 * s[0] + s[1] + s[2]) % P, (s[0] - s[1] + s[2]) % P, (s[0] + s[1] - 2 * s[2]) % P]        # params.rs:15-19

    def half(state, keys):
        rounds = []
        for rk in keys:
            state = [(s + k) % P for s, k in zip(state, rk)]
            rounds.append(list(state))
            state = mds([pow(s, 3, P) for s in state])
        return rounds, state

    first, state = half([v % P for v in inputs], params["full1"])
    partial = []
    for rk in params["partial_opt"]:
        state[2] = (state[2] + rk) % P
        partial.append(state[2])
        state[2] = pow(state[2], 3, P)
        state = mds(state)
    keys2 = [list(k) for k in params["full2"]]
    keys2[0] = [2841653098167170594677968593255398661749780759922623066311132183067080032372,
                3013664908435951456462052676857400233978317167153628607151632758126998548956,
                1580909581709481477620907470438960344056357690169203419381231226301063390430]                       # poseidon/mod.rs:100-104
    second, out = half(state, keys2)
    # cross-check against the plain permutation (poseidon/mod.rs:166-197)
    s, rnd = [v % P for v in inputs], 0
    allk = params["full1"] + params["partial"] + params["full2"]
    for _ in range(4):
        s = mds([pow((a + k) % P, 3, P) for a, k in zip(s, allk[rnd])]); rnd += 1
    for _ in range(83):
        s = [(a + k) % P for a, k in zip(s, allk[rnd])]; s[2] = pow(s[2], 3, P); s = mds(s); rnd += 1
    for _ in range(4):
        s = mds([pow((a + k) % P, 3, P) for a, k in zip(s, allk[rnd])]); rnd += 1
    assert s == out, "optimised Poseidon schedule differs from the plain permutation"
    _poseidon_cache[key] = dict(full=first + second, partial=partial, out=out)
    return _poseidon_cache[key]


def load_poseidon_params(path):
    d = json.load(open(path))
    h = lambda v: int(v, 16)
    return dict(full1=[[h(x) for x in r] for r in d["FULL_ROUND_KEYS_1ST_HALF"]], full2=[[h(x) for x in r] for r in d["FULL_ROUND_KEYS_2ND_HALF"]],
                partial=[[h(x) for x in r] for r in d["PARTIAL_ROUND_KEYS"]], partial_opt=[h(x) for x in d["PARTIAL_ROUND_KEYS_OPTIMIZED"]])


class StarknetTrace:
    """ExecutionTrace::new of the starknet layout (layouts/src/starknet/trace.rs:98-995).  private_input: "pedersen" [(index, a, b)],
    "range_check" [(index, value)], "bitwise" [(index, x, y)]; the ECDSA / EC-op / Poseidon builtins are filled with the reference's
    dummy instances (the committed bootloader fixture has none of its own)."""

    def __init__(self, public_input: AirPublicInput, register_states, memory, poseidon_params, private_input=None):
        priv = private_input or {}
        priv = {k: priv.get(k, []) for k in ("pedersen", "range_check", "bitwise")}
        self.public_input = public_input
        num_cycles = len(register_states)
        assert num_cycles & (num_cycles - 1) == 0
        n = self.trace_len = num_cycles * 16
        self.public_memory = list(public_input.public_memory)
        padding = self.padding_entry = public_input.public_memory_padding()
        seg = public_input.memory_segments
        words = {}

        def word_at(pc):
            if pc not in words:
                words[pc] = Word(memory[pc])
            return words[pc]

        flags = [0] * n
        npc = [padding[0], padding[1]] * (n // 2)
        rc_pool = []
        for ap, fp, pc in register_states:
            w = word_at(pc)
            rc_pool += [w.get_off_dst(), w.get_off_op0(), w.get_off_op1()]
        rc128 = [(idx, val, RecursiveTrace._rc_parts(val)) for idx, val in priv["range_check"]]
        for _, _, parts in rc128:
            rc_pool += parts
        ordered_rc, rc_padding = ordered_with_padding(rc_pool)
        self.range_check_min, self.range_check_max = min(rc_pool), max(rc_pool)
        rc_max = self.range_check_max
        ordered_rc, rc_padding = iter(ordered_rc), iter(rc_padding)
        rc_col = [rc_max] * n
        aux = [0] * n
        N, A = SN_NPC, SN_AUX
        for c, (ap, fp, pc) in enumerate(register_states):                     # trace.rs:186-251
            w, o = word_at(pc), c * 16
            op0, op1, dst = w.get_op0(ap, fp, memory), w.get_op1(pc, ap, fp, memory), w.get_dst(ap, fp, memory)
            for f in range(16):
                flags[o + f] = w.get_flag_prefix(f)
            npc[o + N["Pc"]], npc[o + N["Instruction"]] = pc, w.w
            npc[o + N["MemOp0Addr"]], npc[o + N["MemOp0"]] = w.get_op0_addr(ap, fp), op0
            npc[o + N["MemDstAddr"]], npc[o + N["MemDst"]] = w.get_dst_addr(ap, fp), dst
            npc[o + N["MemOp1Addr"]], npc[o + N["MemOp1"]] = w.get_op1_addr(pc, ap, fp, memory), op1
            for off in range(0, 16, SN["PUB_MEM_STEP"]):
                npc[o + off + N["PubMemAddr"]] = npc[o + off + N["PubMemVal"]] = 0
            rc_col[o + RANGE_CHECK["OffDst"]], rc_col[o + RANGE_CHECK["OffOp1"]], rc_col[o + RANGE_CHECK["OffOp0"]] = \
                w.get_off_dst(), w.get_off_op1(), w.get_off_op0()
            aux[o + A["Tmp0"]], aux[o + A["Tmp1"]] = w.get_tmp0(ap, fp, memory), w.get_tmp1(pc, ap, fp, memory)
            aux[o + A["Ap"]], aux[o + A["Fp"]] = ap, fp
            aux[o + A["Op0MulOp1"]], aux[o + A["Res"]] = op0 * op1 % P, w.get_res(pc, ap, fp, memory)
        for index in range(len(rc128), num_cycles // SN["RC_RATIO"]):          # :253-270
            value = 0
            for _ in range(8):
                value = (value << 16) + next(rc_padding, rc_max)
            rc128.append((index, value, RecursiveTrace._rc_parts(value)))
        for cycle in range(num_cycles):                                        # :272-298
            o = cycle * 16
            if cycle % 2 == 1:
                rc_col[o + RANGE_CHECK["Unused"]] = next(rc_padding, rc_max)
            for off in range(0, 16, RANGE_CHECK_STEP):
                rc_col[o + off + RANGE_CHECK["Ordered"]] = next(ordered_rc, rc_max)
        assert next(rc_padding, None) is None and next(ordered_rc, None) is None
        for o in range(0, n, SN["DILUTED_STEP"]):                              # diluted cells share the column: clear them (:304-313)
            rc_col[o + 1] = rc_col[o + 5] = 0
        # Pedersen: one step per row, 512 rows per hash, its own four columns (:315-387)
        ped_x, ped_y, ped_suffix, ped_slope = [0] * n, [0] * n, [0] * n, [0] * n
        ped_begin = seg["pedersen"][0]
        instances = list(priv["pedersen"])
        for k in range(n // 512):
            index, a, b = instances[k] if k < len(instances) else (k, 0, 0)
            t, base = pedersen_instance_trace(a % P, b % P), k * 512
            for s, (pt, suffix, slope) in enumerate(t["steps"]):
                ped_x[base + s], ped_y[base + s] = pt
                ped_suffix[base + s], ped_slope[base + s] = suffix % P, slope
            (a3, a2), (b3, b2) = t["a_flags"], t["b_flags"]
            ped_slope[base + 255], ped_slope[base + 256 + 255] = a2, b2          # Pedersen::Bit251AndBit196 = 255 (column 4)
            aux[base + 71], aux[base + 256 + 71] = a3, b3                          # Bit251AndBit196AndBit192 = 71 (column 8)
            addr = ped_begin + index * 3
            npc[base + N["PedersenInput0Addr"]], npc[base + N["PedersenInput0Addr"] + 1] = addr, a % P
            npc[base + N["PedersenInput1Addr"]], npc[base + N["PedersenInput1Addr"] + 1] = addr + 1, b % P
            npc[base + N["PedersenOutputAddr"]], npc[base + N["PedersenOutputAddr"] + 1] = addr + 2, t["output"]
        # range-check builtin: 256 rows per value (:389-425)
        rc_begin = seg["range_check"][0]
        for k in range(n // 256):
            index, value, parts = rc128[k]
            base = k * 256
            for j, part in enumerate(parts):
                rc_col[base + RC_BUILTIN_COMPONENT + 32 * j] = part
            npc[base + N["RangeCheck128Addr"]], npc[base + N["RangeCheck128Addr"] + 1] = rc_begin + index, value % P
        # ECDSA: 32768 rows per signature, all dummies (:427-525)
        E = SN_ECDSA
        ec_begin = seg["ecdsa"][0]
        t = ecdsa_dummy_trace()
        for k in range(n // 32768):
            base = k * 32768
            for half, (steps, dbl) in enumerate(((t["rq_steps"], t["pubkey_doubling"]), (t["wb_steps"], t["b_doubling"]))):
                for i in range(256):
                    r = base + 64 * (256 * half + i)
                    (partial, _, suffix, slope, xdi), (dpt, dslope) = steps[i], dbl[i]
                    aux[r + E["PubkeyDoublingX"]], aux[r + E["PubkeyDoublingY"]], aux[r + E["PubkeyDoublingSlope"]] = dpt[0], dpt[1], dslope
                    aux[r + E["PubkeyPartialSumX"]], aux[r + E["PubkeyPartialSumY"]] = partial
                    aux[r + E["PubkeyPartialSumSlope"]], aux[r + E["PubkeyPartialSumXDiffInv"]], aux[r + E["RSuffix"]] = slope, xdi, suffix % P
            for i in range(256):
                r = base + 128 * i
                partial, _, suffix, slope, xdi = t["zg_steps"][i]
                aux[r + E["GeneratorPartialSumX"]], aux[r + E["GeneratorPartialSumY"]] = partial
                aux[r + E["GeneratorPartialSumSlope"]], aux[r + E["GeneratorPartialSumXDiffInv"]], aux[r + E["MessageSuffix"]] = slope, xdi, suffix % P
            for name, key in (("BSlope", "b_slope"), ("BXDiffInv", "b_x_diff_inv"), ("WInv", "w_inv"), ("RInv", "r_inv"), ("RPointSlope", "r_point_slope"),
                              ("RPointXDiffInv", "r_point_x_diff_inv"), ("MessageInv", "message_inv")):
                aux[base + E[name]] = t[key]
            aux[base + E["PubkeyXSquared"]] = t["pubkey"][0] ** 2 % P
            npc[base + N["EcdsaPubkeyAddr"]], npc[base + N["EcdsaPubkeyAddr"] + 1] = ec_begin + 2 * k, t["pubkey"][0]
            npc[base + N["EcdsaMessageAddr"]], npc[base + N["EcdsaMessageAddr"] + 1] = ec_begin + 2 * k + 1, t["message"]
        # bitwise: 1024 rows per instance; its diluted chunks live in the range-check column (:527-649)
        bw_begin = seg["bitwise"][0]
        bw_instances, diluted_pool, dummy = list(priv["bitwise"]), [], None
        for k in range(n // 1024):
            index, x, y = bw_instances[k] if k < len(bw_instances) else (k, 0, 0)
            base = k * 1024
            if (x, y) == (0, 0) and dummy is not None:
                diluted_pool += dummy
            else:
                before = len(diluted_pool)
                parts = [partition256(v) for v in (x, y, x & y, x ^ y)]
                v = [parts[2][3][j] + parts[3][3][j] for j in range(4)]
                for j, sh in enumerate((4, 4, 4, 8)):
                    s = v[j] << sh
                    diluted_pool.append(undilute(s))
                    rc_col[base + SN_BITWISE_SHIFTED[j]] = s
                for q, part in enumerate(parts):
                    for chunk in range(4):
                        for j in range(4):
                            rc_col[base + 256 * q + 1 + 16 * (4 * chunk + j)] = part[chunk][j]
                            diluted_pool.append(undilute(part[chunk][j]))
                if (x, y) == (0, 0):
                    dummy = diluted_pool[before:]
            o = base + N["BitwisePoolAddr"]
            for j, val in enumerate((x, y, x & y, x ^ y)):
                npc[o + 256 * j], npc[o + 256 * j + 1] = bw_begin + index * 5 + j, val % P
            npc[base + N["BitwiseXOrYAddr"]], npc[base + N["BitwiseXOrYAddr"] + 1] = bw_begin + index * 5 + 4, (x | y) % P
        ordered_dil, dil_padding = ordered_with_padding(diluted_pool, 0, (1 << DILUTED_CHECK_N_BITS) - 1)
        ordered_dil, dil_padding = [dilute(v) for v in ordered_dil], iter(dilute(v) for v in dil_padding)
        done = False                                                            # padding into the odd diluted steps (:666-693)
        for base in range(0, n, 1024):
            for i in range(1, 128, 2):
                off = i * 8 + 1
                if off in SN_BITWISE_SHIFTED:
                    continue
                v = next(dil_padding, None)
                if v is None:
                    done = True
                    break
                rc_col[base + off] = v
            if done:
                break
        steps = n // 8
        for k, v in enumerate(ordered_dil):                                     # :695-701
            rc_col[8 * (steps - len(ordered_dil) + k) + 5] = v
        assert next(dil_padding, None) is None
        # EC-op: 16384 rows per instance, all dummies (:708-777)
        O = SN_ECOP
        eo_begin = seg["ec_op"][0]
        t = ec_op_dummy_trace()
        for k in range(n // 16384):
            base = k * 16384
            for i in range(256):
                r = base + 64 * i
                (dpt, dslope), (partial, _, suffix, slope, xdi) = t["q_doubling"][i], t["r_steps"][i]
                aux[r + O["QDoublingX"]], aux[r + O["QDoublingY"]], aux[r + O["QDoublingSlope"]] = dpt[0], dpt[1], dslope
                aux[r + O["RPartialSumX"]], aux[r + O["RPartialSumY"]], aux[r + O["MSuffix"]] = partial[0], partial[1], suffix % P
                if i != 255:
                    aux[r + O["RPartialSumSlope"]], aux[r + O["RPartialSumXDiffInv"]] = slope, xdi
            aux[base + O["MBit251AndBit196"]] = aux[base + O["MBit251AndBit196AndBit192"]] = 0
            for j, (name, val) in enumerate((("EcOpPXAddr", t["p"][0]), ("EcOpPYAddr", t["p"][1]), ("EcOpQXAddr", t["q"][0]), ("EcOpQYAddr", t["q"][1]),
                                             ("EcOpMAddr", t["m"]), ("EcOpRXAddr", t["r"][0]), ("EcOpRYAddr", t["r"][1]))):
                npc[base + N[name]], npc[base + N[name] + 1] = eo_begin + 7 * k + j, val
        # Poseidon: 512 rows per instance, all-zero inputs (:779-873)
        Q = SN_POSEIDON
        po_begin = seg["poseidon"][0]
        t = poseidon_trace((0, 0, 0), poseidon_params)
        for k in range(n // 512):
            base = k * 512
            for i, st in enumerate(t["full"]):
                r = base + 64 * i
                for j, (cell, sq) in enumerate((("Full0", "Full0Sq"), ("Full1", "Full1Sq"), ("Full2", "Full2Sq"))):
                    aux[r + Q[cell]], aux[r + Q[sq]] = st[j], st[j] * st[j] % P
            for i, v in enumerate(t["partial"][:64]):
                rc_col[base + 8 * i + Q["Partial0"]], rc_col[base + 8 * i + Q["Partial0Sq"]] = v, v * v % P
            for i, v in enumerate(t["partial"][61:]):
                aux[base + 16 * i + Q["Partial1"]], aux[base + 16 * i + Q["Partial1Sq"]] = v, v * v % P
            for j, (name, val) in enumerate((("PoseidonInput0Addr", 0), ("PoseidonInput1Addr", 0), ("PoseidonInput2Addr", 0),
                                             ("PoseidonOutput0Addr", t["out"][0]), ("PoseidonOutput1Addr", t["out"][1]), ("PoseidonOutput2Addr", t["out"][2]))):
                npc[base + N[name]], npc[base + N[name] + 1] = po_begin + 6 * k + j, val
        # memory (:875-929)
        accesses = [(npc[i], npc[i + 1]) for i in range(0, n, 2)]
        srt = sorted(accesses + self.public_memory, key=lambda e: e[0])
        gaps = []
        for (a, _), (b, _) in zip(srt, srt[1:]):
            gaps.extend(range(a + 1, b))
        gaps = iter(gaps)
        for o in range(0, n, 16):
            addr = next(gaps, None)
            if addr is None:
                break
            npc[o + N["UnusedAddr"]], npc[o + N["UnusedVal"]] = addr, 0
        assert next(gaps, None) is None
        accesses = [(npc[i], npc[i + 1]) for i in range(0, n, 2)]
        cells = n // SN["PUB_MEM_STEP"]
        ordered = sorted(accesses + [padding] * (cells - len(self.public_memory)) + self.public_memory, key=lambda e: e[0])
        zeros, ordered = ordered[:cells], ordered[cells:]
        assert all(a == 0 for a, _ in zeros) and ordered[0][0] == 1
        for cur, nxt in zip(ordered, ordered[1:]):
            assert cur == nxt or cur[0] == nxt[0] - 1, (cur, nxt)
        memory_col = [v for e in ordered for v in e]
        assert len(memory_col) == n
        self.base_columns = [flags, ped_x, ped_y, ped_suffix, ped_slope, npc, memory_col, rc_col, aux]        # trace.rs:931-941
        self.initial_registers, self.final_registers = register_states[0], register_states[-1]

    def build_extension_columns(self, challenges):
        """starknet/trace.rs:997-1100: ONE extension column."""
        n = self.trace_len
        npc, mem, rc = self.base_columns[5], self.base_columns[6], self.base_columns[7]

        def running_quotient(num_factors, den_factors):
            nums, dens, a, b = [], [], 1, 1
            for f, g in zip(num_factors, den_factors):
                a, b = a * f % P, b * g % P
                nums.append(a)
                dens.append(b)
            return [x * y % P for x, y in zip(nums, batch_inverse(dens))], a, b

        perm = [0] * n
        z, alpha = challenges[MEM_Z], challenges[MEM_A]
        perm[0::2], _, _ = running_quotient(((z - (alpha * npc[i + 1] + npc[i])) % P for i in range(0, n, 2)),
                                            ((z - (alpha * mem[i + 1] + mem[i])) % P for i in range(0, n, 2)))
        z = challenges[RC_Z]
        perm[1::4], a, b = running_quotient(((z - rc[i]) % P for i in range(0, n, 4)), ((z - rc[i + 2]) % P for i in range(0, n, 4)))
        assert a == b
        z = challenges[DILUTED_PERM_Z]
        perm[7::8], a, b = running_quotient(((z - rc[i + 1]) % P for i in range(0, n, 8)), ((z - rc[i + 5]) % P for i in range(0, n, 8)))
        assert a == b
        z, alpha = challenges[DILUTED_AGG_Z], challenges[DILUTED_AGG_A]
        acc = 1
        perm[3] = 1
        for i in range(1, n // 8):
            u = (rc[8 * i + 5] - rc[8 * (i - 1) + 5]) % P
            acc = (acc * (1 + z * u) + alpha * u * u) % P
            perm[8 * i + 3] = acc
        return [perm]

    def gen_hints(self, challenges):
        """starknet AirConfig::gen_hints (starknet/air.rs:2408-2476), PublicInputHint order (:3244-3262)."""
        pi = self.public_input
        s, k = self.trace_len // SN["PUB_MEM_STEP"], len(self.public_memory)
        z, alpha = challenges[MEM_Z], challenges[MEM_A]
        den = 1
        for a, v in self.public_memory:
            den = den * (z - (alpha * v + a)) % P
        pad = pow((z - (alpha * self.padding_entry[1] + self.padding_entry[0])) % P, s - k, P)
        quotient = pow(z, s, P) * pow(den * pad % P, -1, P) % P
        cumulative = compute_diluted_cumulative_value(challenges[DILUTED_AGG_Z], challenges[DILUTED_AGG_A])
        seg = pi.memory_segments
        return [pi.initial_ap(), pi.initial_pc(), pi.final_ap(), pi.final_pc(), quotient, 1, pi.rc_min, pi.rc_max, 1, 0, cumulative,
                seg["pedersen"][0], seg["range_check"][0], seg["ecdsa"][0], seg["bitwise"][0], seg["ec_op"][0], seg["poseidon"][0]]


def load_bootloader(directory, poseidon_params_path):
    """The reference's example/bootloader fixture (starknet layout, 131072 steps; trace.bin is kept gzipped in tests/golden)."""
    import gzip
    import os
    import tempfile

    pub = AirPublicInput.from_file(os.path.join(directory, "air-public-input.json"))
    with tempfile.NamedTemporaryFile() as f:
        f.write(gzip.open(os.path.join(directory, "trace.bin.gz")).read())
        f.flush()
        regs = read_register_states(f.name)
    mem = read_memory(os.path.join(directory, "memory.bin"))
    priv = json.load(open(os.path.join(directory, "air-private-input.json")))
    private = {"pedersen": [(e["index"], int(e["x"], 16), int(e["y"], 16)) for e in priv.get("pedersen", [])],
               "range_check": [(e["index"], int(e["value"], 16)) for e in priv.get("range_check", [])],
               "bitwise": [(e["index"], int(e["x"], 16), int(e["y"], 16)) for e in priv.get("bitwise", [])]}
    return StarknetTrace(pub, regs, mem, load_poseidon_params(poseidon_params_path), private)
