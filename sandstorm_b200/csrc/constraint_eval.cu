// Composition-constraint evaluation over the LDE domain: executes the
// straight-line program produced by sandstorm_b200/air/program.py — the flattened Expr DAG of
// AirConfig::composition_constraint (layouts/src/recursive/air.rs:1184-1200) — on every LDE row.
//
// One thread per row i (x_i = 3 * w_N^i).  Row-periodic sub-expressions (zerofiers and their inverses,
// periodic columns) arrive as lookup tables indexed by i mod T; trace taps read
// lde[col][(i + offset*blowup) mod N] — neighbouring threads read neighbouring elements of the same
// column, so every tap is a coalesced 32-byte-per-lane stream served mostly by L2 (each LDE element
// is touched once per tap offset).  Full-period denominators (X - g^e boundary terms) are inverted
// together with one batched inversion per row.  Values live in a per-thread slot file.
//
// Algorithmic bytes per row: (C_base + C_ext + 1) * 32 B; field-ops per row are reported by the compiler
// (CompiledProgram.n_mul / n_addsub).
#include "ctx.h"
#include "pedersen.cuh"   // ec::inv_chain

using namespace ss;

namespace {

enum Op : uint32_t { OP_NOP, OP_CONST, OP_TRACE, OP_TABLE, OP_X, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_INV, OP_BATCHINV, OP_OUT, OP_MULC, OP_ADDC };
// two instantiations: constraint compositions need ~32 slots and a handful of boundary denominators;
// DEEP compositions pin one denominator per out-of-domain point (191 + 1 for the starknet layout)
constexpr int MAX_SLOTS = 256;
constexpr int MAX_BATCH = 192;
constexpr int SMALL_SLOTS = 64;
constexpr int SMALL_BATCH = 32;
constexpr uint32_t MAGIC = 0x50435353u;
constexpr int T_XLO = 20, T_XHI = 21;

struct EvalArgs {
    const uint4 *code;
    int n_instr;
    const Fp *consts;
    const Fp *tables;
    const uint2 *tdesc;        // (log_period, offset)
    const Fp *cols;
    unsigned long long stride;
    int log_N;
    const Fp *xlo, *xhi;       // 3 * w_N^i (i < 4096), w_N^(4096 i)
    Fp *out;
    unsigned long long row_begin, row_count;   // rows [row_begin, row_begin + row_count) of the LDE domain
};

__device__ __forceinline__ Fp ldg_fp(const Fp *p) {
    const uint4 *q = reinterpret_cast<const uint4 *>(p);
    const uint4 a = __ldg(q), b = __ldg(q + 1);
    Fp v;
    v.l[0] = a.x; v.l[1] = a.y; v.l[2] = a.z; v.l[3] = a.w; v.l[4] = b.x; v.l[5] = b.y; v.l[6] = b.z; v.l[7] = b.w;
    return v;
}

template <int SLOTS, int BATCH>
__global__ void __launch_bounds__(128) constraint_eval_kernel(const EvalArgs A) {
    const unsigned long long t = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
    const unsigned long long N = 1ull << A.log_N;
    if (t >= A.row_count) return;
    const unsigned long long i = A.row_begin + t;
    Fp s[SLOTS];
#pragma unroll 1
    for (int pc = 0; pc < A.n_instr; ++pc) {
        const uint4 ins = __ldg(A.code + pc);
        const uint32_t op = ins.x & 0xffu, d = ins.x >> 8, a = ins.y, b = ins.z;
        switch (op) {
        case OP_CONST: s[d] = ldg_fp(A.consts + a); break;
        case OP_TRACE: {
            const long long off = (long long)(int)ins.w;
            const unsigned long long row = (unsigned long long)((long long)i + off) & (N - 1);
            s[d] = ldg_fp(A.cols + (unsigned long long)a * A.stride + row);
            break;
        }
        case OP_TABLE: {
            const uint2 td = __ldg(A.tdesc + a);
            s[d] = ldg_fp(A.tables + td.y + (i & ((1ull << td.x) - 1)));
            break;
        }
        case OP_X: {
            Fp v = ldg_fp(A.xlo + (i & 4095ull));
            if (i >> 12) v = fp::mul(v, ldg_fp(A.xhi + (i >> 12)));
            s[d] = v;
            break;
        }
        case OP_ADD: s[d] = fp::add(s[a], s[b]); break;
        case OP_SUB: s[d] = fp::sub(s[a], s[b]); break;
        case OP_MUL: s[d] = fp::mul(s[a], s[b]); break;
        case OP_MULC: s[d] = fp::mul(s[a], ldg_fp(A.consts + b)); break;
        case OP_ADDC: s[d] = fp::add(s[a], ldg_fp(A.consts + b)); break;
        case OP_NEG: s[d] = fp::neg(s[a]); break;
        case OP_INV: s[d] = ec::inv_chain(s[a]); break;
        case OP_BATCHINV: {
            // Montgomery's trick over slots [a, a + b)
            Fp pre[BATCH];
            Fp acc = fp::one();
            for (uint32_t k = 0; k < b; ++k) { pre[k] = acc; acc = fp::mul(acc, s[a + k]); }
            Fp inv = ec::inv_chain(acc);
            for (uint32_t k = b; k-- > 0;) {
                const Fp t = fp::mul(inv, pre[k]);
                inv = fp::mul(inv, s[a + k]);
                s[a + k] = t;
            }
            break;
        }
        case OP_OUT: {
            const Fp v = fp::canon(s[a]);
            uint4 *q = reinterpret_cast<uint4 *>(A.out + i);
            q[0] = make_uint4(v.l[0], v.l[1], v.l[2], v.l[3]);
            q[1] = make_uint4(v.l[4], v.l[5], v.l[6], v.l[7]);
            break;
        }
        default: break;
        }
    }
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
                             void *d_out, void *stream) {
    if (!ctx) return SS_ERR_INVALID;
    if (!h_program || program_bytes < 32 || !d_lde_cols || !d_out || n_cols < 1)
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: bad arguments");
    const uint32_t *w = static_cast<const uint32_t *>(h_program);
    if (w[0] != MAGIC || w[1] != 1) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: not a program blob (magic/version)");
    const uint32_t n_instr = w[2], n_consts = w[3], n_tables = w[4], n_slots = w[5];
    if ((int)w[6] != log_n || (int)w[7] != log_blowup)
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: program compiled for log_n=%u blowup=%u, called with %d/%d", w[6], w[7], log_n, log_blowup);
    if (n_slots > (uint32_t)MAX_SLOTS) return fail(ctx, SS_ERR_UNSUPPORTED, "ss_constraint_eval: %u slots > %d", n_slots, MAX_SLOTS);
    const int log_N = log_n + log_blowup;
    if (col_stride < (1ull << log_N)) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: col_stride too small");
    const size_t n_tdesc = (size_t)n_tables + (n_tables & 1u);          // descriptor area padded to 16 bytes
    size_t head_words = 8 + 2 * n_tdesc + 4 * (size_t)n_instr;
    size_t head_bytes = (head_words * 4 + 31) / 32 * 32;
    size_t table_elems = 0;
    for (uint32_t t = 0; t < n_tables; ++t) {
        const uint32_t lp = w[8 + 2 * t], off = w[8 + 2 * t + 1];
        if (lp > 20 || off != table_elems) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: corrupt table descriptor %u", t);
        table_elems += (size_t)1 << lp;
    }
    if (program_bytes != head_bytes + 32 * ((size_t)n_consts + table_elems))
        return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: blob size mismatch");
    // validate operands so that a bad program cannot index outside the slot file / matrix
    const uint32_t *code = w + 8 + 2 * n_tdesc;
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const uint32_t op = code[4 * pc] & 0xff, d = code[4 * pc] >> 8, a = code[4 * pc + 1], b = code[4 * pc + 2];
        bool ok = d < n_slots || op == OP_OUT || op == OP_BATCHINV || op == OP_NOP;
        switch (op) {
        case OP_CONST: ok = ok && a < n_consts; break;
        case OP_TRACE: ok = ok && a < (uint32_t)n_cols; break;
        case OP_TABLE: ok = ok && a < n_tables; break;
        case OP_ADD: case OP_SUB: case OP_MUL: ok = ok && a < n_slots && b < n_slots; break;
        case OP_MULC: case OP_ADDC: ok = ok && a < n_slots && b < n_consts; break;
        case OP_NEG: case OP_INV: case OP_OUT: ok = ok && a < n_slots; break;
        case OP_BATCHINV: ok = a + b <= n_slots && b <= (uint32_t)MAX_BATCH; break;
        case OP_X: case OP_NOP: break;
        default: ok = false;
        }
        if (!ok) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: invalid instruction %u (op %u)", pc, op);
    }
    SS_CUDA_CHECK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = pick_stream(ctx, stream);
    // upload the program (stream-ordered; freed after the kernel on the same stream)
    uint8_t *d_prog = nullptr;
    SS_CUDA_CHECK(ctx, cudaMallocAsync(reinterpret_cast<void **>(&d_prog), program_bytes, st));
    SS_CUDA_CHECK(ctx, cudaMemcpyAsync(d_prog, h_program, program_bytes, cudaMemcpyHostToDevice, st));
    Fp *xlo, *xhi;
    const size_t N = (size_t)1 << log_N;
    ss_status rc;
    if ((rc = cached_table(ctx, {T_XLO, log_N, 0}, N < 4096 ? N : 4096, fill_xlo, &xlo))) return rc;
    if ((rc = cached_table(ctx, {T_XHI, log_N, 0}, N <= 4096 ? 1 : N / 4096, fill_xhi, &xhi))) return rc;
    EvalArgs A;
    A.tdesc = reinterpret_cast<const uint2 *>(d_prog + 32);
    A.code = reinterpret_cast<const uint4 *>(d_prog + 32 + 8 * n_tdesc);
    A.n_instr = (int)n_instr;
    A.consts = reinterpret_cast<const Fp *>(d_prog + head_bytes);
    A.tables = A.consts + n_consts;
    A.cols = static_cast<const Fp *>(d_lde_cols);
    A.stride = col_stride;
    A.log_N = log_N;
    A.xlo = xlo; A.xhi = xhi;
    A.out = static_cast<Fp *>(d_out);
    if (row_count == 0) { row_begin = 0; row_count = N; }                 // 0 = the whole domain
    if (row_begin + row_count > N) return fail(ctx, SS_ERR_INVALID, "ss_constraint_eval: row range outside the domain");
    A.row_begin = row_begin; A.row_count = row_count;
    uint32_t max_batch = 0;
    for (uint32_t pc = 0; pc < n_instr; ++pc)
        if ((code[4 * pc] & 0xff) == OP_BATCHINV && code[4 * pc + 2] > max_batch) max_batch = code[4 * pc + 2];
    if (n_slots <= (uint32_t)SMALL_SLOTS && max_batch <= (uint32_t)SMALL_BATCH)
        constraint_eval_kernel<SMALL_SLOTS, SMALL_BATCH><<<(unsigned)((row_count + 127) / 128), 128, 0, st>>>(A);
    else
        constraint_eval_kernel<MAX_SLOTS, MAX_BATCH><<<(unsigned)((row_count + 127) / 128), 128, 0, st>>>(A);
    ctx->launches++;
    SS_CUDA_CHECK(ctx, cudaGetLastError());
    SS_CUDA_CHECK(ctx, cudaFreeAsync(d_prog, st));
    // the host blob may be pageable: make sure the copy has consumed it before returning
    SS_CUDA_CHECK(ctx, cudaStreamSynchronize(st));
    return SS_OK;
}

}  // extern "C"
