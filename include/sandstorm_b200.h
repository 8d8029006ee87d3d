/* sandstorm_b200 — C ABI of the B200 (sm_100a) proving backend.
 *
 * This is the drop-in boundary for the ONE hot path of `sandstorm prove`:
 * per-column LDE, Cairo-AIR constraint-composition evaluation, Merkle commitment, FRI folding,
 * DEEP composition.  The reference (Rust) has no FFI of its own: the seam is the trait surface
 * between sandstorm and ministark.  Each entry point below names the reference interface it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` block and trait impls that bind it.
 *
 * Conventions
 *   - every function returns ss_status (0 = OK, < 0 = error; ss_last_error(ctx) has the text).
 *     Nothing panics or throws across the ABI (the reference panics via unwrap(), e.g.
 *     crypto/src/merkle/mod.rs:116,120 — the Rust shim maps a non-zero status to that panic).
 *   - Fp252 element = 4 x u64 little-endian limbs of x*2^256 mod p, canonical (< p): the exact
 *     bytes of ark-ff's Fp256<MontBackend<_,4>> that reference crypto/src/utils.rs:15-17 exposes.
 *     Goldilocks element = 1 x u64 (see ss_field).
 *   - matrices are column-major, one contiguous column of 2^log_rows elements every `col_stride`
 *     elements: the layout of ministark::Matrix (Vec<GpuVec<F>>), SURVEY.md §8 a1.
 *   - d_* pointers are device memory of the ctx's GPU, h_* pointers are host memory.
 *   - `stream` is a cudaStream_t passed as void* (NULL = the CUDA default stream).  Calls are
 *     asynchronous on that stream unless stated; one ss_ctx per (thread, device).
 */
#ifndef SANDSTORM_B200_H
#define SANDSTORM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ss_ctx ss_ctx;
typedef struct ss_tree ss_tree;

typedef enum {
    SS_OK = 0,
    SS_ERR_INVALID = -1,      /* bad argument                                   */
    SS_ERR_CUDA = -2,         /* CUDA runtime / launch failure                  */
    SS_ERR_OOM = -3,          /* device allocation failed                       */
    SS_ERR_UNSUPPORTED = -4   /* valid request this build does not implement    */
} ss_status;

typedef enum {
    SS_FIELD_FP252 = 0,       /* p = 2^251 + 17*2^192 + 1  (cli/src/main.rs:25-26)  */
    SS_FIELD_GOLDILOCKS = 1   /* p = 2^64 - 2^32 + 1       (cli/src/main.rs:30)     */
} ss_field;

typedef enum {
    SS_ORDER_NATURAL = 0,
    SS_ORDER_BITREV = 1,        /* rows in bit-reversed order: what ministark commits (pinned by the reference's proof artefacts) */
    SS_ORDER_BITREV_RC = 3      /* ss_merkle_build / ss_hash_rows only: rows AND columns bit-reversed — a FRI layer: leaf r holds the
                                   `fold` evaluations that fold together, as consecutive entries of the bit-reversed vector */
} ss_order;

/* Which reference MatrixMerkleTree the commitment reproduces (src/claims.rs:10-32). */
typedef enum {
    SS_TREE_KECCAK = 0,        /* LeafVariantMerkleTree<Keccak256HashFn>            recursive::EthVerifierClaim  */
    SS_TREE_KECCAK_M20 = 1,    /* LeafVariantMerkleTree<MaskedKeccak256HashFn<20>>  starknet::EthVerifierClaim   */
    SS_TREE_FRIENDLY = 2,      /* FriendlyMerkleTree<N, PedersenHashFn>             *::CairoVerifierClaim        */
    SS_TREE_BLAKE2S_M20 = 3,   /* masked Blake2s at every level (Friendly with N = 0)                            */
    SS_TREE_SHA256 = 4         /* ministark MatrixMerkleTreeImpl<Sha256HashFn> (Goldilocks claims)               */
} ss_tree_kind;

/* ------------------------------------------------------------------ context / memory */
int ss_version(void);
ss_status ss_create(int device, ss_ctx **out);
void ss_destroy(ss_ctx *ctx);
const char *ss_last_error(const ss_ctx *ctx);
ss_status ss_sync(ss_ctx *ctx);
/* number of CUDA kernels this ctx has launched so far (bench.py reports the per-step delta) */
uint64_t ss_kernel_launches(const ss_ctx *ctx);
/* tuning switches and read-back values of a context (no reference counterpart).  Keys:
 *   "ce_aot"      1 (default) = ss_constraint_eval may use a kernel specialised at build time, 0 = always interpret
 *   "ce_minb"     resident CTAs per SM the interpreter is compiled for (5, 6 or 7)
 *   "ce_last_aot" (read) 1 when the last ss_constraint_eval call ran a specialised kernel */
ss_status ss_set_option(ss_ctx *ctx, const char *key, int64_t value);
int64_t ss_get_option(const ss_ctx *ctx, const char *key, int64_t dflt);
/* thin wrappers so that a host without the CUDA runtime (Rust) can own device buffers;
 * replaces ministark_gpu's GpuAllocator / GpuVec role (layouts/src/recursive/trace.rs:55-56,115) */
ss_status ss_malloc(ss_ctx *ctx, size_t bytes, void **d_ptr);
ss_status ss_free(ss_ctx *ctx, void *d_ptr);
ss_status ss_host_register(ss_ctx *ctx, void *h_ptr, size_t bytes);     /* pin a Vec for async copies */
ss_status ss_host_unregister(ss_ctx *ctx, void *h_ptr);
ss_status ss_memcpy_h2d(ss_ctx *ctx, void *d_dst, const void *h_src, size_t bytes, void *stream);
ss_status ss_memcpy_d2h(ss_ctx *ctx, void *h_dst, const void *d_src, size_t bytes, void *stream);

/* ------------------------------------------------------------------ NTT / LDE (§8 a2, a3, a8)
 * ss_ntt: per-column ark-poly Radix2EvaluationDomain::{fft, ifft, coset_fft, coset_ifft}
 * (what ministark Matrix::interpolate / Matrix::evaluate run; domain generator 3^((p-1)/n),
 * pinned by builtins/src/pedersen/periodic.rs:1184-1209; coset offset = Fp::GENERATOR = 3).
 * In place on d_cols.  inverse includes the 1/n factor.  in/out_order let a caller skip the
 * bit-reversal permutation (DIF: natural->bitrev, DIT: bitrev->natural are the native forms). */
ss_status ss_ntt(ss_ctx *ctx, ss_field field, void *d_cols, uint64_t col_stride, int n_cols, int log_n,
                 int inverse, int coset, ss_order in_order, ss_order out_order, void *stream);

/* ss_lde: Matrix::interpolate(trace_domain) then Matrix::evaluate(lde_domain) fused:
 * trace evaluations on <w_n>  ->  evaluations on 3*<w_N>, N = n << log_blowup.
 * d_coeffs (optional, n per column, stride coeff_stride) receives the interpolated polynomials
 * (coset-scaled: c_k * 3^k, BIT-REVERSED order) — the operand the DEEP / OOD stage consumes. */
ss_status ss_lde(ss_ctx *ctx, ss_field field, const void *d_trace, uint64_t trace_stride, int n_cols,
                 int log_n, int log_blowup, void *d_lde, uint64_t lde_stride, void *d_coeffs,
                 uint64_t coeff_stride, ss_order out_order, void *stream);

/* ss_coset_eval: the polynomials whose coefficients ss_lde left in d_coeffs (coefficient k at position brev(k), scaled or
 * not) evaluated on the coset h<w_n>: dst[j] = sum_k coeff_k h^k w_n^(j k), natural order.  With ss_lde's c_k 3^k and
 * h = z/3 this yields T(z g^j) for every j: all out-of-domain mask values of a column (ministark's OOD evaluation of
 * air.trace_arguments()) from one transform.  d_dst may equal d_coeffs. */
ss_status ss_coset_eval(ss_ctx *ctx, ss_field field, const void *d_coeffs, uint64_t coeff_stride, int n_cols, int log_n,
                        const void *h_h, void *d_dst, uint64_t dst_stride, void *stream);

/* ------------------------------------------------------------------ Merkle (§8 a9-a13)
 * ss_merkle_build = MatrixMerkleTree::from_matrix (crypto/src/merkle/mod.rs:110-123, :289-304):
 * n_cols == 1 -> raw-leaf variant, n_cols >= 2 -> row hashes (crypto/src/merkle/utils.rs:19-46)
 * then MerkleTreeImpl::new.  row_order = SS_ORDER_BITREV commits row brev(i) as leaf i (fused
 * bit-reversal).  n_friendly = N_FRIENDLY_LAYERS (22 in src/claims.rs:10), SS_TREE_FRIENDLY only.
 * The tree stays on the device until ss_tree_free (trees must outlive prove(), §8b ownership). */
ss_status ss_merkle_build(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_cols,
                          uint64_t col_stride, int n_cols, int log_rows, ss_order row_order,
                          ss_tree **out, void *stream);
/* Pieces of ss_merkle_build for commitments split over several GPUs: the digests of the tree leaves
 * [leaf_begin, leaf_begin + leaf_count) of the matrix (leaf p = row brev(p) for the bit-reversed orders) written to
 * d_out[k]; a tree over leaf digests computed elsewhere; the in-place bit-reversal permutation of 32-byte items. */
ss_status ss_hash_rows(ss_ctx *ctx, ss_tree_kind kind, const void *d_cols, uint64_t col_stride, int n_cols, int log_rows,
                       ss_order order, uint64_t leaf_begin, uint64_t leaf_count, void *d_out, void *stream);
ss_status ss_merkle_build_from_leaves(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_leaf_digests, int log_rows,
                                      ss_tree **out, void *stream);
ss_status ss_bitrev_permute32(ss_ctx *ctx, void *d_items, int log_n, void *stream);
/* MerkleTree::root -> Digest::as_bytes (32 bytes).  Synchronises. */
ss_status ss_merkle_root(ss_ctx *ctx, const ss_tree *tree, uint8_t root[32]);
/* Copies node i (1 = root .. 2^log_rows - 1; storage form: byte digest, or Montgomery limbs for
 * algebraic levels) / leaf digests for MerkleTree::prove path extraction.  Synchronises. */
ss_status ss_merkle_nodes(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n,
                          uint8_t *h_out /* n * 32 */);
ss_status ss_merkle_leaves(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n,
                           uint8_t *h_out /* n * 32 */);
/* MerkleTree::prove(indices): sibling path of each index, leaf level first (log_rows * 32 bytes each) */
ss_status ss_merkle_open(ss_ctx *ctx, const ss_tree *tree, const uint64_t *h_indices, size_t n,
                         uint8_t *h_paths /* n * log_rows * 32 */);
/* Root of a row-sharded commitment: 2^log_count sub-tree roots (one per GPU row range, in row
 * order, as ss_merkle_root returns them) -> root of the whole tree.  For SS_TREE_FRIENDLY the combined
 * levels are Pedersen (log_count <= N_FRIENDLY) and each sub-tree is built with n_friendly - log_count.
 * Synchronises. */
ss_status ss_merkle_combine(ss_ctx *ctx, ss_tree_kind kind, const uint8_t *h_subroots, int log_count, uint8_t root[32]);
/* The upper log_count siblings of the authentication path of any leaf under sub-tree `index`, bottom first, in
 * ss_merkle_open's storage form: ss_merkle_open(sub-tree, local leaf) followed by these is the path through the whole
 * tree (MerkleTree::prove, crypto/src/merkle/mod.rs:120-148).  subroots_algebraic: the sub-roots are Pedersen felts
 * (SS_TREE_FRIENDLY with log_count < n_friendly, or a single-column tree).  Synchronises. */
ss_status ss_merkle_combine_open(ss_ctx *ctx, ss_tree_kind kind, const uint8_t *h_subroots, int log_count, int subroots_algebraic,
                                 uint64_t index, uint8_t *h_path /* log_count * 32 */);
int ss_tree_log_rows(const ss_tree *tree);
void ss_tree_free(ss_tree *tree);
/* batch Pedersen hash (builtins/src/pedersen/mod.rs:31-36), Montgomery limbs in and out */
ss_status ss_pedersen_hash(ss_ctx *ctx, const void *d_a, const void *d_b, void *d_out, size_t n, void *stream);
/* Matrix::read_row for the query phase: h_rows[i*n_cols + j] = cols[j][idx[i]].  Synchronises. */
ss_status ss_rows_gather(ss_ctx *ctx, const void *d_cols, uint64_t col_stride, int n_cols,
                         const uint64_t *h_indices, size_t n, void *h_rows);

/* ------------------------------------------------------------------ FRI / DEEP (§8 a14, a15) */
/* One FRI fold by F = 2^log_fold (1..4) of evaluations on h*<w_N>, natural order (ministark
 * FriProver::build_layers / apply_drp, fold factor from cli/src/main.rs:57-58):
 *   f(x) = sum_{j<F} x^j F_j(x^F)   ->   out[i] = sum_j alpha^j F_j(x_i^F),   i < N/F.
 * flags bit 0: multiply by F (StarkWare's binary folds with alpha, alpha^2, alpha^4 and no 1/2).
 * The layer's commitment is ss_merkle_build(d_evals, col_stride = N/F, n_cols = F, log_rows = log_n - log_fold):
 * row i of the reference's FRI layer matrix is (e[i], e[i + N/F], ...), i.e. the buffer itself.
 * out_begin / out_count select the outputs [out_begin, out_begin + out_count) (count 0 = all): the
 * row range one GPU folds when a layer is sharded by rows; d_out is indexed by the absolute output row. */
ss_status ss_fri_fold(ss_ctx *ctx, ss_field field, const void *d_evals, int log_n, int log_fold,
                      const void *h_alpha, const void *h_domain_offset, int flags, uint64_t out_begin,
                      uint64_t out_count, void *d_out, void *stream);
/* d_out[i] = 1 / (x_i - c) for every point x_i = 3 * w_N^i of the LDE coset (N = 2^log_n), one batched
 * inversion per 16 rows.  Because x_i - c g^e = g^e (x_{i - b e} - c) for the trace-domain generator
 * g = w_N^b, every boundary denominator X - g^e of the AIR (c = 1) and every DEEP denominator
 * X - z g^e (c = z) is a shifted read of such a vector. */
ss_status ss_inv_x_minus_c(ss_ctx *ctx, ss_field field, int log_n, int log_row_step /* only rows that are multiples of 2^step are written */,
                           uint64_t row_begin, uint64_t row_count /* in units of 2^step rows, wrapping mod N; 0 = all */,
                           const void *h_c, void *d_out, void *stream);
/* Out-of-domain evaluations (trace / composition polynomials at z * g^k): for e < n_evals,
 * h_out[e] = poly_{h_cols[e]}(h_points[e]).  natural_order = 0: the coefficient matrix ss_lde writes
 * (coset-scaled, bit-reversed); 1: plain coefficients in natural order (composition columns).  Synchronises. */
ss_status ss_poly_eval(ss_ctx *ctx, ss_field field, const void *d_coeffs, uint64_t coeff_stride, int log_n, int natural_order,
                       const int32_t *h_cols, const void *h_points, size_t n_evals, void *h_out);

/* The same out-of-domain values computed from the TRACE evaluations (barycentric form; no coefficient vectors):
 * h_out[e] = T_{h_cols[e]}(z * g^h_offsets[e]) with g the generator of the trace domain and d_trace_cols the
 * column-major trace (n = 2^log_n rows, natural order) — the mask of air.trace_arguments() (ministark).
 * row_count != 0 restricts the sum to the trace rows [row_begin, row_begin + row_count): the results of disjoint
 * ranges add up to the value (one range per GPU).  z must lie outside the trace domain.  Synchronises. */
ss_status ss_ood_eval(ss_ctx *ctx, ss_field field, const void *d_trace_cols, uint64_t col_stride, int log_n,
                      const int32_t *h_cols, const uint64_t *h_offsets, size_t n_evals, const void *h_z,
                      uint64_t row_begin, uint64_t row_count, void *h_out);

/* ------------------------------------------------------------------ proof of work (§8 f.3)
 * PublicCoin::grind_proof_of_work (crypto/src/public_coin/solidity.rs:120-141 with hash_kind 0 = Keccak-256,
 * cairo.rs:133-154 with hash_kind 1 = Blake2s-256): the SMALLEST nonce >= 1 whose hash
 * H(H(0x0123456789ABCDED || digest || bits) || nonce_be) starts with `bits` zero bits.  Synchronises. */
ss_status ss_pow_grind(ss_ctx *ctx, int hash_kind, const uint8_t digest[32], int bits, uint64_t *nonce_out);

/* ------------------------------------------------------------------ constraint evaluation (§8 a4-a7)
 * Evaluates a compiled composition-constraint program (the Expr DAG of
 * AirConfig::composition_constraint, layouts/src/recursive/air.rs:1184-1200, flattened by the host
 * into straight-line code; format in sandstorm_b200/air/program.py) on every LDE row — or on every
 * 2^log_row_step-th row: a quotient of degree < n such as the DEEP composition is fixed by its n values on
 * the sub-coset of the rows that are multiples of the blowup, and is extended with ss_ntt afterwards.   */
ss_status ss_constraint_eval(ss_ctx *ctx, const void *h_program, size_t program_bytes,
                             const void *d_lde_cols, uint64_t col_stride, int n_cols, int log_n,
                             int log_blowup, uint64_t row_begin, uint64_t row_count /* 0 = all rows */,
                             int log_row_step /* rows row_begin + k * 2^log_row_step, k < row_count */,
                             void *d_out /* indexed by (absolute row >> log_row_step) */, void *stream);

/* ------------------------------------------------------------------ row-sharded transforms over W GPUs (§8e plan B)
 * A transform of size M over W ranks = one all-to-all between a size-W transform across the ranks (ss_shard_dft) and
 * local size-M/W transforms (ss_ntt_shard); sandstorm_b200/parallel.py drives them with torch.distributed / NCCL.
 * ss_ntt_shard: stages & 1 = inverse DIF of size 2^log_m (natural -> bit-reversed), coefficient k scaled by c0 * h0^k;
 *               stages & 2 = forward DIT of size 2^(log_m + log_expand) from bit-reversed coefficients (zero-padded),
 *               natural output k scaled by tw^k (h_tw NULL: none).  stages = 3: both (the local LDE).
 * ss_shard_dft: out[k1 * out_stride + i] = tw(i)^k1 * sum_{j1<W} in[j1 * in_stride + i] * w_W^(+-j1 k1), i < count,
 *               tw(i) = w_(2^tw_log_m)^(+-(tw_offset + i)) (tw_log_m < 0: none); W = 2^log_w <= 8. */
ss_status ss_ntt_shard(ss_ctx *ctx, ss_field field, const void *d_src, uint64_t src_stride, int n_cols, int log_m, int stages,
                       int log_expand, const void *h_c0, const void *h_h0, const void *h_tw, void *d_dst, uint64_t dst_stride,
                       void *stream);
ss_status ss_shard_dft(ss_ctx *ctx, ss_field field, const void *d_in, uint64_t in_stride, void *d_out, uint64_t out_stride,
                       uint64_t count, int log_w, int inverse, int tw_log_m, uint64_t tw_offset, void *stream);

/* Collectives of the row-sharded hot path over NCCL, one context per process / device (csrc/dist.cu; NCCL is bound at run time
 * with dlopen, so single-GPU users need no NCCL).  Vectors and matrices are full-length device buffers of which this rank owns
 * the block-cyclic pieces: rows [k m + r s, k m + (r+1) s), k < W, m = rows / W, s = m / W.
 *   ss_dist_unique_id   rank 0 creates the 128-byte NCCL id; the host passes it to the other ranks by its own means
 *   ss_dist_init        joins the communicator (collective); world = 2, 4 or 8
 *   ss_dist_lde         one column: evaluations on <w_n> (src_on_coset: on 3<w_n>) -> evaluations on 3<w_N>, both block-cyclic
 *   ss_dist_halo        the `halo` rows that follow every owned piece of every column, from the next rank (halo <= s)
 *   ss_dist_commit      MatrixMerkleTree::from_matrix over all ranks' rows in bit-reversed leaf order -> root (collective, synchronises);
 *                       out_subtree (nullable): this rank's sub-tree over leaves [rank N/W, (rank+1) N/W), kept for ss_dist_open and
 *                       freed by the caller; out_subroots (nullable, 32 W bytes): every rank's sub-root
 *   ss_dist_allgather   in-place all-gather of a block-cyclic vector
 *   ss_dist_open        MerkleTree::prove for n leaf positions of the whole tree (collective: the owner of a leaf range opens its
 *                       sub-tree, the upper siblings come from the sub-roots; every rank receives every path, n * log_rows * 32 bytes)
 *   ss_dist_gather_rows Matrix::read_row for n natural row indices of a block-cyclic matrix (collective; every rank receives every row) */
ss_status ss_dist_unique_id(ss_ctx *ctx, uint8_t id[128]);
ss_status ss_dist_init(ss_ctx *ctx, const uint8_t id[128], int rank, int world);
ss_status ss_dist_finalize(ss_ctx *ctx);
ss_status ss_dist_lde(ss_ctx *ctx, ss_field field, const void *d_src, int log_n, int log_blowup, int src_on_coset, void *d_dst,
                      void *stream);
ss_status ss_dist_halo(ss_ctx *ctx, void *d_cols, uint64_t col_stride, int n_cols, int log_rows, uint64_t halo, void *stream);
ss_status ss_dist_allgather(ss_ctx *ctx, void *d_vec, int log_rows, void *stream);
ss_status ss_dist_commit(ss_ctx *ctx, ss_tree_kind kind, int n_friendly, const void *d_cols, uint64_t col_stride, int n_cols,
                         int log_rows, uint8_t root[32], ss_tree **out_subtree, uint8_t *out_subroots, void *stream);
ss_status ss_dist_open(ss_ctx *ctx, ss_tree_kind kind, int subroots_algebraic, const ss_tree *subtree, const uint8_t *h_subroots,
                       const uint64_t *h_positions, size_t n, uint8_t *h_paths, void *stream);
ss_status ss_dist_gather_rows(ss_ctx *ctx, const void *d_cols, uint64_t col_stride, int n_cols, int log_rows, const uint64_t *h_indices,
                              size_t n, void *h_rows, void *stream);

/* ------------------------------------------------------------------ extension columns (§8 f1)
 * Trace::build_extension_columns (layouts/src/recursive/trace.rs:699-814, starknet/trace.rs:997-1100) as device prefix
 * scans instead of the reference's sequential loops + batch_inversion.  Elements are read at index j * stride of the
 * given column pointers (the layouts interleave virtual columns: memory pairs at stride 2, range-check cells at
 * stride 4, ...) and result i is written at d_out[i * out_stride].
 *   ss_perm_product:      out[i] = prod_{j<=i} (z - (alpha * num_v[j] + num_a[j])) / (z - (alpha * den_v[j] + den_a[j]))
 *                         (d_num_v = d_den_v = NULL: z - num_a[j] over z - den_a[j])          trace.rs:706-755
 *   ss_diluted_aggregate: out[0] = 1, out[i] = out[i-1] * (1 + z u_i) + alpha u_i^2, u_i = d[i] - d[i-1]   trace.rs:787-806 */
ss_status ss_perm_product(ss_ctx *ctx, ss_field field, const void *d_num_a, const void *d_num_v, const void *d_den_a,
                          const void *d_den_v, uint64_t stride, uint64_t count, const void *h_z, const void *h_alpha,
                          void *d_out, uint64_t out_stride, void *stream);
ss_status ss_diluted_aggregate(ss_ctx *ctx, ss_field field, const void *d_ordered, uint64_t stride, uint64_t count,
                               const void *h_z, const void *h_alpha, void *d_out, uint64_t out_stride, void *stream);

#ifdef __cplusplus
}
#endif
#endif
