// Composition-constraint evaluation over the LDE domain: executes the
// straight-line program produced by sandstorm_b200/air/program.py — the flattened Expr DAG of
// AirConfig::composition_constraint (layouts/src/recursive/air.rs:1184-1200) — on every LDE row.
//
// One thread per row i (x_i = 3 * w_N^i).  Operands are fetched where they are used: a slot of the
// per-thread value file, a constant, an entry of a row-periodic lookup table (zerofiers and their
// inverses, periodic columns; index i mod T) or a trace tap lde[col][(i + offset) mod N] —
// neighbouring threads read neighbouring elements of the same column, so every tap is a coalesced
// 32-byte-per-lane stream served mostly by L1/L2 (each LDE element is touched once per tap offset).
//
// The compiler knows an exact upper bound of every value, so the arithmetic here is branch-free: ADD is a
// raw 256-bit addition, SUBK adds k * p before subtracting, RED is emitted only where a bound would
// overflow, and a linear combination with general coefficients is one DOT: its products are accumulated
// unreduced in 512 bits (fp::WideAcc) and Montgomery-reduced once.
//
// Algorithmic bytes per row: (C_base + C_ext + 1) * 32 B; field-ops per row are reported by the compiler
// (CompiledProgram.n_mul / n_addsub).
#include "ctx.h"
#include "pedersen.cuh"   // ec::inv_chain
#include <cstdlib>

using namespace ss;

namespace {

enum Op : uint32_t { OP_NOP, OP_MOV, OP_ADD, OP_SUBK, OP_RED, OP_MUL, OP_DOT, OP_INV, OP_OUT, OP_COUNT };
enum Kind : uint32_t { K_SLOT, K_CONST, K_TAP, K_TABLE, K_X, K_COUNT };
// two instantiations of the slot file: the Cairo compositions and DEEP quotients need < 32 slots
constexpr int MAX_SLOTS = 256;
constexpr int SMALL_SLOTS = 32;
constexpr uint32_t MAGIC = 0x50435353u;
constexpr uint32_t VERSION = 3;
constexpr int T_XLO = 20, T_XHI = 21;

struct EvalArgs {
    const uint4 *code;
    int n_words;
    const Fp *consts;
    const Fp *tables;
    const uint2 *tdesc;        // (log_period, offset)
    const uint2 *taps;         // (column, row offset mod N)
    const Fp *cols;
    unsigned long long stride;
    int log_N;
    const Fp *xlo, *xhi;       // 3 * w_N^i (i < 4096), w_N^(4096 i)
    Fp *out;
    unsigned long long row_begin, row_count;   // rows row_begin + (k << log_step), k < row_count, of the LDE domain
    int log_step;
};

__device__ __forceinline__ Fp ldg_fp(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}

// operand word: kind << 29 | payload (program.py).  `w` is warp-uniform, so every branch here is uniform.
__device__ __forceinline__ Fp fetch(const uint32_t w, const Fp *s, const EvalArgs &A, const unsigned long long i) {
    const uint32_t pay = w & 0x1fffffffu;
    switch (w >> 29) {
    case K_SLOT: return s[pay];
    case K_CONST: return ldg_fp(A.consts + pay);
    case K_TAP: {
        const uint2 tp = __ldg(A.taps + pay);
        const unsigned long long row = (i + tp.y) & ((1ull << A.log_N) - 1);
        return ldg_fp(A.cols + (unsigned long long)tp.x * A.stride + row);
    }
    case K_TABLE: {
        const uint2 td = __ldg(A.tdesc + pay);
        return ldg_fp(A.tables + td.y + (i & ((1ull << (td.x & 0xffu)) - 1)));
    }
    default: {                                                                   // K_X
        Fp v = ldg_fp(A.xlo + (i & 4095ull));
        if (i >> 12) v = fp::mul(v, ldg_fp(A.xhi + (i >> 12)));
        return v;
    }
    }
}

// Tables stored as (challenge-independent values, index of a per-proof constant): multiply them out once, in place,
// right after the blob is uploaded (program.py: descriptor word 0 = log_period | (scale + 1) << 8).
__global__ void table_scale_kernel(Fp *tables, const uint2 *tdesc, const Fp *consts, int n_tables) {
    for (int t = 0; t < n_tables; ++t) {
        const uint2 td = tdesc[t];
        if (!(td.x >> 8)) continue;
        const Fp c = consts[(td.x >> 8) - 1];
        const unsigned long long T = 1ull << (td.x & 0xffu);
        for (unsigned long long j = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; j < T; j += (unsigned long long)gridDim.x * blockDim.x)
            tables[td.y + j] = fp::canon(fp::mul(tables[td.y + j], c));
    }
}

// MINB = resident CTAs per SM the register allocation is sized for (5: 96 registers, no spills; 6: 80; 7: 72)
template <int SLOTS, int MINB>
__global__ void __launch_bounds__(128, MINB) constraint_eval_kernel(const EvalArgs A) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    if (t >= A.row_count) return;
    const unsigned long long i = A.row_begin + (t << A.log_step);
    Fp s[SLOTS];
#pragma unroll 1
    for (int pc = 0; pc < A.n_words; ++pc) {
        const uint4 ins = __ldg(A.code + pc);
        const uint32_t op = ins.x & 0xffu, d = (ins.x >> 8) & 0xffu, n = ins.x >> 16;
        switch (op) {
        case OP_MOV: s[d] = fetch(ins.y, s, A, i); break;
        case OP_ADD: s[d] = fp::add_raw(fetch(ins.y, s, A, i), fetch(ins.z, s, A, i)); break;
        case OP_SUBK: s[d] = fp::sub_kp(fetch(ins.y, s, A, i), fetch(ins.z, s, A, i), n); break;
        case OP_RED: s[d] = fp::red(s[d]); break;
        case OP_MUL: s[d] = fp::mul(fetch(ins.y, s, A, i), fetch(ins.z, s, A, i)); break;
        case OP_DOT: {
            fp::WideAcc acc;
            fp::acc_init(acc);
#pragma unroll 1
            for (uint32_t k = 0; k < n; k += 2) {
                const uint4 pr = __ldg(A.code + (++pc));
                fp::acc_mac(acc, fetch(pr.x, s, A, i), fetch(pr.y, s, A, i));
                if (k + 1 < n) fp::acc_mac(acc, fetch(pr.z, s, A, i), fetch(pr.w, s, A, i));
            }
            s[d] = fp::acc_reduce(acc);
            break;
        }
        case OP_INV: s[d] = ec::inv_chain(fetch(ins.y, s, A, i)); break;
        case OP_OUT: {
            const Fp v = fp::canon(fetch(ins.y, s, A, i));
            uint4 *q = reinterpret_cast<uint4 *>(A.out + (i >> A.log_step));
            q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
            q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
            break;
        }
        default: break;
        }
    }
}

// ---- build-time specialisations (tools/gen_ce_kernels.py): the same programs as straight-line code --------
// Run-time values the generated code indexes with literals; passed as a __grid_constant__ kernel parameter,
// i.e. they live in the constant bank and cost no load instruction.
struct GenArgs {
    uint32_t tap_off[1024];    // row offset of tap k (mod N)
    uint32_t tab_off[128];     // element offset of table t inside the table area
};
struct Wide { uint32_t t[16]; };

// One shared copy of the 8x8-limb product / reduction: the straight-line callers stay short (instruction
// cache), and ptxas keeps arguments and results in registers (no stack traffic — checked in the SASS).
__device__ __noinline__ Fp mul_ni(Fp a, Fp b) { return fp::mul(a, b); }
__device__ __noinline__ Wide wide_p_ni(Fp a, Fp b) { Wide w; fp::mul_wide_plus_p<true>(w.t, a, b); return w; }
__device__ __noinline__ Wide wide_ni(Fp a, Fp b) { Wide w; fp::mul_wide_plus_p<false>(w.t, a, b); return w; }
__device__ __noinline__ Fp reduce_ni(Wide w) { return fp::mont_reduce(w.t); }
__device__ __forceinline__ void wide_add(Wide &acc, const Wide &w) {
    using namespace ptx;
    acc.t[0] = add_cc(acc.t[0], w.t[0]);
#pragma unroll
    for (int k = 1; k < 15; ++k) acc.t[k] = addc_cc(acc.t[k], w.t[k]);
    acc.t[15] = addc(acc.t[15], w.t[15]);
}
__device__ __forceinline__ Fp fetch_x(const EvalArgs &A, const unsigned long long i) {
    Fp v = ldg_fp(A.xlo + (i & 4095ull));
    if (i >> 12) v = mul_ni(v, ldg_fp(A.xhi + (i >> 12)));
    return v;
}
__device__ __forceinline__ void store_out(Fp *dst, const Fp &v) {
    uint4 *q = reinterpret_cast<uint4 *>(dst);
    q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
    q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
}

#include "ce_gen.cuh"

// 64-bit FNV-1a over the structural part of the blob — the mirror of program.py structure_hash()
unsigned long long structure_hash(const uint32_t *w, size_t n_tdesc, size_t n_tapd) {
    const uint32_t n_words = w[2], n_tables = w[4], n_taps = w[8];
    unsigned long long h = 0xCBF29CE484222325ull;
    auto mix = [&](uint32_t v) {
        for (int sh = 0; sh < 32; sh += 8) h = (h ^ ((v >> sh) & 0xffu)) * 0x100000001B3ull;
    };
    mix(n_words); mix(w[3]); mix(n_tables); mix(w[5]); mix(w[7]); mix(n_taps);
    const uint32_t *tdesc = w + 16, *tapd = tdesc + 2 * n_tdesc, *code = tapd + 2 * n_tapd;
    for (uint32_t t = 0; t < n_tables; ++t) mix(tdesc[2 * t]);
    for (uint32_t t = 0; t < n_taps; ++t) mix(tapd[2 * t]);
    for (size_t k = 0; k < 4 * (size_t)n_words; ++k) mix(code[k]);
    return h;
}

Fp host_root(int log_n) {
    uint32_t e[8] = {0, 0, 0, 0, 0, 0, 0x00000011u, 0x08000000u};
    for (int s = 0; s < log_n; ++s)
        for (int i = 0; i < 8; ++i) { e[i] >>= 1; if (i < 7) e[i] |= e[i + 1] << 31; }
    return fp::canon(fp::pow_limbs(fp::from_u32(3), e, 8));
}
void fill_xlo(Fp *dst, size_t n, int log_n, int) {
    const Fp w = host_root(log_n);
    Fp c = fp::from_u32(3);
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}
void fill_xhi(Fp *dst, size_t n, int log_n, int) {
    const Fp w = fp::pow_u64(host_root(log_n), 4096);
    Fp c = fp::one();
    for (size_t i = 0; i < n; ++i) { dst[i] = fp::canon(c); c = fp::mul(c, w); }
}

}  // namespace

extern "C" {

ss_status ss_constraint_eval(ss_ctx *ctx, const void *h_program, size_t program_bytes, const void *d_lde_cols,
                             uint64_t col_stride, int n_cols, int log_n, int log_blowup, uint64_t row_begin, uint64_t row_count,
                             int log_row_step, void *d_out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (!h_program || program_bytes < 64 || !d_lde_cols || !d_out || n_cols < 1)
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: bad arguments");
    const uint32_t *w = static_cast<const uint32_t *>(h_program);
    if (w[0] != MAGIC || w[1] != VERSION) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: not a program blob (magic/version)");
    const uint32_t n_words = w[2], n_consts = w[3], n_tables = w[4], n_slots = w[5], n_taps = w[8];
    if ((int)w[6] != log_n || (int)w[7] != log_blowup)
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: program compiled for log_n=%u blowup=%u, called with %d/%d", w[6], w[7], log_n, log_blowup);
    if (n_slots > (uint32_t)MAX_SLOTS) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_constraint_eval: %u slots > %d", n_slots, MAX_SLOTS);
    const int log_N = log_n + log_blowup;
    if (log_N > 32) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: domain too large");
    if (col_stride < (1ull << log_N)) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: col_stride too small");
    const size_t n_tdesc = (size_t)n_tables + (n_tables & 1u);          // descriptor area padded to 16 bytes
    const size_t n_tapd = (size_t)n_taps + (n_taps & 1u);
    size_t head_words = 16 + 2 * n_tdesc + 2 * n_tapd + 4 * (size_t)n_words;
    size_t head_bytes = (head_words * 4 + 31) / 32 * 32;
    if (head_bytes > program_bytes) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: blob size mismatch");
    size_t table_elems = 0;
    bool scaled_tables = false;
    for (uint32_t t = 0; t < n_tables; ++t) {
        const uint32_t lp = w[16 + 2 * t] & 0xffu, scale = w[16 + 2 * t] >> 8, off = w[16 + 2 * t + 1];
        if (lp > 20 || off != table_elems || scale > n_consts) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: corrupt table descriptor %u", t);
        table_elems += (size_t)1 << lp;
        scaled_tables |= scale != 0;
    }
    if (program_bytes != head_bytes + 32 * ((size_t)n_consts + table_elems))
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: blob size mismatch");
    // validate every operand so that a bad program cannot index outside the slot file, the tables or the matrix
    const uint32_t *tapd = w + 16 + 2 * n_tdesc;
    for (uint32_t t = 0; t < n_taps; ++t)
        if (tapd[2 * t] >= (uint32_t)n_cols || (uint64_t)tapd[2 * t + 1] >= (1ull << log_N))
            return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: tap %u outside the matrix", t);
    const uint32_t *code = tapd + 2 * n_tapd;
    auto operand_ok = [&](uint32_t word) {
        const uint32_t pay = word & 0x1fffffffu;
        switch (word >> 29) {
        case K_SLOT: return pay < n_slots;
        case K_CONST: return pay < n_consts;
        case K_TAP: return pay < n_taps;
        case K_TABLE: return pay < n_tables;
        case K_X: return true;
        default: return false;
        }
    };
    for (uint32_t pc = 0; pc < n_words; ++pc) {
        const uint32_t w0 = code[4 * pc], op = w0 & 0xff, d = (w0 >> 8) & 0xff, n = w0 >> 16;
        bool ok = d < n_slots || op == OP_OUT || op == OP_NOP;
        switch (op) {
        case OP_NOP: break;
        case OP_MOV: case OP_INV: case OP_OUT: ok = ok && operand_ok(code[4 * pc + 1]); break;
        case OP_ADD: case OP_MUL: ok = ok && operand_ok(code[4 * pc + 1]) && operand_ok(code[4 * pc + 2]); break;
        case OP_SUBK: ok = ok && n <= 31 && operand_ok(code[4 * pc + 1]) && operand_ok(code[4 * pc + 2]); break;
        case OP_RED: break;
        case OP_DOT: {
            const uint32_t extra = (n + 1) / 2;
            ok = ok && n >= 1 && pc + extra < n_words;
            for (uint32_t k = 0; ok && k < n; ++k) {
                const uint32_t *pr = code + 4 * (pc + 1 + k / 2) + 2 * (k & 1);
                ok = operand_ok(pr[0]) && operand_ok(pr[1]);
            }
            if (ok) pc += extra;
            break;
        }
        default: ok = false;
        }
        if (!ok) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: invalid instruction at word %u (op %u)", pc, op);
    }
    if (log_row_step < 0 || log_row_step > log_N) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: bad row step");
    const size_t N = (size_t)1 << log_N;
    if (row_count == 0) { row_begin = 0; row_count = N >> log_row_step; }   // 0 = the whole domain
    if ((row_begin & ((1ull << log_row_step) - 1)) || row_begin + (row_count << log_row_step) > N)
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: row range outside the domain");
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    Fp *xlo, *xhi;
    ss_status rc;
    if ((rc = cached_table(ctx, {T_XLO, log_N, 0}, N < 4096 ? N : 4096, fill_xlo, &xlo))) return rc;
    if ((rc = cached_table(ctx, {T_XHI, log_N, 0}, N <= 4096 ? 1 : N / 4096, fill_xhi, &xhi))) return rc;
    // upload the program through the context's pinned staging area: stream-ordered, the caller is not blocked and
    // may reuse its blob as soon as this returns.  (Everything that can fail has been checked above: no leak paths.)
    uint8_t *d_prog = nullptr;
    SS_CUDA_CHECK(ctx, dev_alloc(ctx, reinterpret_cast<void **>(&d_prog), program_bytes));
    if ((rc = stage_upload(ctx, d_prog, h_program, program_bytes, st))) { dev_free(ctx, d_prog); return rc; }
    EvalArgs A;
    A.tdesc = reinterpret_cast<const uint2 *>(d_prog + 64);
    A.taps = reinterpret_cast<const uint2 *>(d_prog + 64 + 8 * n_tdesc);
    A.code = reinterpret_cast<const uint4 *>(d_prog + 64 + 8 * n_tdesc + 8 * n_tapd);
    A.n_words = (int)n_words;
    A.consts = reinterpret_cast<const Fp *>(d_prog + head_bytes);
    A.tables = A.consts + n_consts;
    A.cols = static_cast<const Fp *>(d_lde_cols);
    A.stride = col_stride;
    A.log_N = log_N;
    A.xlo = xlo; A.xhi = xhi;
    A.out = static_cast<Fp *>(d_out);
    if (scaled_tables) {
        table_scale_kernel<<<148, 256, 0, st>>>(const_cast<Fp *>(A.tables), A.tdesc, A.consts, (int)n_tables);
        ctx->launches++;
    }
    A.row_begin = row_begin; A.row_count = row_count; A.log_step = log_row_step;
    // tuning switches: ss_set_option, with the environment as the default
    static const int env_minb = [] { const char *e = getenv("SS_CE_MINB"); return e ? atoi(e) : 5; }();
    static const int env_aot = [] { const char *e = getenv("SS_CE_AOT"); return e ? atoi(e) : 1; }();
    const int minb = (int)option(ctx, "ce_minb", env_minb);
    const bool use_gen = option(ctx, "ce_aot", env_aot) != 0;
    const unsigned grid = (unsigned)((row_count + 127) / 128);
    const GenEntry *gen = nullptr;
    if (use_gen && n_taps <= 1024 && n_tables <= 128) {
        const unsigned long long h = structure_hash(w, n_tdesc, n_tapd);
        const int want = (int)option(ctx, "ce_aot_minb", 0);            // 0 = the variant chosen at build time
        for (const GenEntry &e : GEN_KERNELS)
            if (e.hash == h && (want ? e.minb == want : e.dflt)) gen = &e;
    }
    ctx->options["ce_last_aot"] = gen ? 1 : 0;
    if (gen) {
        GenArgs G;
        for (uint32_t t = 0; t < n_taps; ++t) G.tap_off[t] = tapd[2 * t + 1];
        for (uint32_t t = 0; t < n_tables; ++t) G.tab_off[t] = w[16 + 2 * t + 1];
        gen->kernel<<<grid, 128, 0, st>>>(A, G);
    } else if (n_slots > (uint32_t)SMALL_SLOTS)
        constraint_eval_kernel<MAX_SLOTS, 5><<<grid, 128, 0, st>>>(A);
    else if (minb >= 7)
        constraint_eval_kernel<SMALL_SLOTS, 7><<<grid, 128, 0, st>>>(A);
    else if (minb == 6)
        constraint_eval_kernel<SMALL_SLOTS, 6><<<grid, 128, 0, st>>>(A);
    else
        constraint_eval_kernel<SMALL_SLOTS, 5><<<grid, 128, 0, st>>>(A);
    ctx->launches++;
    const cudaError_t launch_err = cudaGetLastError();
    dev_free(ctx, d_prog);                                  // handed out again in stream order (ctx.h)
    SS_CUDA_CHECK(ctx, launch_err);
    return SS_OK;
}

}  // extern "C"
