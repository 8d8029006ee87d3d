// Entry points declared in include/sandstorm_b200.h that this build does not implement yet.
// They fail loudly with SS_ERR_UNSUPPORTED (never a CPU fallback).
#include "ctx.h"
using namespace ss;

extern "C" {

#ifndef SS_HAVE_MERKLE
ss_status ss_merkle_build(ss_ctx *ctx, ss_tree_kind, int, const void *, uint64_t, int, int, ss_order, ss_tree **, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_build: not built"); }
ss_status ss_merkle_root(ss_ctx *ctx, const ss_tree *, uint8_t *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_root: not built"); }
ss_status ss_merkle_nodes(ss_ctx *ctx, const ss_tree *, const uint64_t *, size_t, uint8_t *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_nodes: not built"); }
ss_status ss_merkle_leaves(ss_ctx *ctx, const ss_tree *, const uint64_t *, size_t, uint8_t *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_leaves: not built"); }
ss_status ss_merkle_open(ss_ctx *ctx, const ss_tree *, const uint64_t *, size_t, uint8_t *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_open: not built"); }
ss_status ss_merkle_combine(ss_ctx *ctx, ss_tree_kind, const uint8_t *, int, uint8_t *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_merkle_combine: not built"); }
int ss_tree_log_rows(const ss_tree *) { return -1; }
void ss_tree_free(ss_tree *) {}
ss_status ss_pedersen_hash(ss_ctx *ctx, const void *, const void *, void *, size_t, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_pedersen_hash: not built"); }
ss_status ss_rows_gather(ss_ctx *ctx, const void *, uint64_t, int, const uint64_t *, size_t, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_rows_gather: not built"); }
#endif

#ifndef SS_HAVE_FRI
ss_status ss_inv_x_minus_c(ss_ctx *ctx, ss_field, int, int, uint64_t, uint64_t, const void *, void *, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_inv_x_minus_c: not built"); }
ss_status ss_fri_fold(ss_ctx *ctx, ss_field, const void *, int, int, const void *, const void *, int, uint64_t, uint64_t, void *, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_fri_fold: not built"); }
ss_status ss_poly_eval(ss_ctx *ctx, ss_field, const void *, uint64_t, int, int, const int32_t *, const void *, size_t, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_poly_eval: not built"); }
ss_status ss_ood_eval(ss_ctx *ctx, ss_field, const void *, uint64_t, int, const int32_t *, const uint64_t *, size_t, const void *, uint64_t, uint64_t, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_ood_eval: not built"); }
#endif

#ifndef SS_HAVE_CONSTRAINTS
ss_status ss_constraint_eval(ss_ctx *ctx, const void *, size_t, const void *, uint64_t, int, int, int, uint64_t, uint64_t, int, void *, void *) { return fail(ctx, SS_ERR_UNSUPPORTED, "ss_constraint_eval: not built"); }
#endif

}
