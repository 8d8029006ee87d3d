// sandstorm_b200.hpp — C++ host side above the C ABI (include/sandstorm_b200.h).
//
// The reference is compiled code (Rust); its toolchain is not in the build image, so this header is the compiled-language
// mirror of the trait surface the hot path sits behind: the same names and argument meaning as
//   ministark::Matrix<Fp>::{interpolate, evaluate}            (call sites layouts/src/recursive/air.rs:66-67)
//   MatrixMerkleTree::{from_matrix, root, prove, prove_rows}  (crypto/src/merkle/mod.rs:64-166, :254-347)
// with the reference's error behaviour (construction panics via unwrap(), crypto/src/merkle/mod.rs:116,120,295,301 ->
// here a thrown sandstorm_b200::Error carrying ss_last_error).  It owns nothing but handles: elements stay the 4 x u64
// Montgomery limbs of ark-ff's Fp256 (crypto/src/utils.rs:15-17), matrices stay column-major.  Header-only; link with
// -lsandstorm_b200.  The Rust binding of INTEGRATION.md has the same shape.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "sandstorm_b200.h"

namespace sandstorm_b200 {

struct Error : std::runtime_error {
    ss_status status;
    Error(ss_status s, const std::string &what) : std::runtime_error(what), status(s) {}
};

// One context per (thread, device): ss_create / ss_destroy.
class Context {
  public:
    explicit Context(int device = 0) {
        ss_ctx *c = nullptr;
        const ss_status s = ss_create(device, &c);
        if (s != SS_OK) throw Error(s, "ss_create failed: no usable CUDA device (there is no CPU fallback)");
        ctx_.reset(c, [](ss_ctx *p) { ss_destroy(p); });
    }
    ss_ctx *get() const { return ctx_.get(); }
    void check(ss_status s) const {
        if (s != SS_OK) throw Error(s, std::string("sandstorm_b200: ") + ss_last_error(ctx_.get()));
    }
    void sync() const { check(ss_sync(ctx_.get())); }

  private:
    std::shared_ptr<ss_ctx> ctx_;
};

using Felt = std::array<uint64_t, 4>;             // Fp252, Montgomery form, little-endian limbs
using Digest = std::array<uint8_t, 32>;

// ministark::Matrix<Fp>: column-major, device resident.
class Matrix {
  public:
    Matrix(Context ctx, int n_cols, int log_rows) : ctx_(std::move(ctx)), n_cols_(n_cols), log_rows_(log_rows) {
        void *p = nullptr;
        ctx_.check(ss_malloc(ctx_.get(), bytes(), &p));
        data_.reset(p, [c = ctx_](void *q) { ss_free(c.get(), q); });
    }
    // Matrix::new(columns): upload of host columns (each num_rows() elements)
    static Matrix from_columns(Context ctx, const std::vector<std::vector<Felt>> &cols) {
        if (cols.empty() || cols[0].empty() || (cols[0].size() & (cols[0].size() - 1))) throw Error(SS_ERR_INVALID, "columns must have 2^k rows");
        int log_rows = 0;
        while ((size_t(1) << log_rows) < cols[0].size()) ++log_rows;
        Matrix m(std::move(ctx), (int)cols.size(), log_rows);
        for (size_t j = 0; j < cols.size(); ++j) {
            if (cols[j].size() != cols[0].size()) throw Error(SS_ERR_INVALID, "ragged matrix");
            m.ctx_.check(ss_memcpy_h2d(m.ctx_.get(), m.column(j), cols[j].data(), m.num_rows() * sizeof(Felt), nullptr));
        }
        m.ctx_.sync();
        return m;
    }
    int num_cols() const { return n_cols_; }
    size_t num_rows() const { return size_t(1) << log_rows_; }
    int log_rows() const { return log_rows_; }
    void *data() const { return data_.get(); }
    void *column(size_t j) const { return static_cast<char *>(data_.get()) + j * num_rows() * sizeof(Felt); }
    const Context &context() const { return ctx_; }

    // Matrix::interpolate(trace_domain): evaluations on <w_n> -> coefficients, natural order
    Matrix interpolate() const {
        Matrix out = clone();
        ctx_.check(ss_ntt(ctx_.get(), SS_FIELD_FP252, out.data(), num_rows(), n_cols_, log_rows_, 1, 0, SS_ORDER_NATURAL, SS_ORDER_NATURAL, nullptr));
        return out;
    }
    // Matrix::evaluate(lde_domain): coefficients -> evaluations on 3 * <w_N>, N = n << log_blowup
    Matrix evaluate(int log_blowup) const {
        Matrix out(ctx_, n_cols_, log_rows_ + log_blowup);
        ctx_.check(zero_fill(out));
        for (int j = 0; j < n_cols_; ++j) ctx_.check(copy_column(out.column(j), column(j)));
        ctx_.check(ss_ntt(ctx_.get(), SS_FIELD_FP252, out.data(), out.num_rows(), n_cols_, out.log_rows(), 0, 1, SS_ORDER_NATURAL, SS_ORDER_NATURAL, nullptr));
        return out;
    }
    // interpolate + evaluate fused (what the prover runs): ss_lde
    Matrix lde(int log_blowup) const {
        Matrix out(ctx_, n_cols_, log_rows_ + log_blowup);
        ctx_.check(ss_lde(ctx_.get(), SS_FIELD_FP252, data(), num_rows(), n_cols_, log_rows_, log_blowup, out.data(), out.num_rows(), nullptr, 0,
                          SS_ORDER_NATURAL, nullptr));
        return out;
    }
    // Matrix::read_row for a set of rows (query phase): row-major [indices.size()][num_cols()]
    std::vector<Felt> rows(const std::vector<uint64_t> &indices) const {
        std::vector<Felt> out(indices.size() * n_cols_);
        ctx_.check(ss_rows_gather(ctx_.get(), data(), num_rows(), n_cols_, indices.data(), indices.size(), out.data()));
        return out;
    }
    std::vector<Felt> column_to_host(size_t j) const {
        std::vector<Felt> out(num_rows());
        ctx_.check(ss_memcpy_d2h(ctx_.get(), out.data(), column(j), bytes() / n_cols_, nullptr));
        ctx_.sync();
        return out;
    }
    Matrix clone() const {
        Matrix out(ctx_, n_cols_, log_rows_);
        for (int j = 0; j < n_cols_; ++j) ctx_.check(copy_column(out.column(j), column(j)));
        return out;
    }

  private:
    size_t bytes() const { return size_t(n_cols_) * num_rows() * sizeof(Felt); }
    // device-to-device helpers through the host-visible ABI (a staging copy: this mirror is for wiring and tests, the
    // prover itself keeps everything resident and never copies whole matrices)
    ss_status copy_column(void *dst, const void *src) const {
        std::vector<Felt> tmp(size_t(1) << log_rows_);
        ss_status s = ss_memcpy_d2h(ctx_.get(), tmp.data(), src, tmp.size() * sizeof(Felt), nullptr);
        if (s == SS_OK) s = ss_sync(ctx_.get());
        if (s == SS_OK) s = ss_memcpy_h2d(ctx_.get(), dst, tmp.data(), tmp.size() * sizeof(Felt), nullptr);
        if (s == SS_OK) s = ss_sync(ctx_.get());
        return s;
    }
    static ss_status zero_fill(const Matrix &m) {
        std::vector<Felt> zeros(m.num_rows(), Felt{0, 0, 0, 0});
        for (int j = 0; j < m.n_cols_; ++j) {
            const ss_status s = ss_memcpy_h2d(m.ctx_.get(), m.column(j), zeros.data(), zeros.size() * sizeof(Felt), nullptr);
            if (s != SS_OK) return s;
        }
        return ss_sync(m.ctx_.get());
    }
    Context ctx_;
    int n_cols_, log_rows_;
    std::shared_ptr<void> data_;
};

// MatrixMerkleTree (crypto/src/merkle/mod.rs): LeafVariantMerkleTree<H> / FriendlyMerkleTree<N, H> by `kind`.
class MatrixMerkleTree {
  public:
    static constexpr int NUM_FRIENDLY_COMMITMENT_LAYERS = 22;   // src/claims.rs:10

    // MatrixMerkleTree::from_matrix(&matrix)
    static MatrixMerkleTree from_matrix(const Matrix &m, ss_tree_kind kind, int n_friendly = NUM_FRIENDLY_COMMITMENT_LAYERS) {
        ss_tree *t = nullptr;
        m.context().check(ss_merkle_build(m.context().get(), kind, kind == SS_TREE_FRIENDLY ? n_friendly : 0, m.data(), m.num_rows(), m.num_cols(),
                                          m.log_rows(), SS_ORDER_NATURAL, &t, nullptr));
        return MatrixMerkleTree(m, t);
    }
    // MerkleTree::root
    Digest root() const {
        Digest d;
        matrix_.context().check(ss_merkle_root(matrix_.context().get(), tree_.get(), d.data()));
        return d;
    }
    // MerkleTree::prove(indices): sibling paths, leaf level first
    std::vector<Digest> prove(const std::vector<uint64_t> &indices) const {
        std::vector<Digest> out(indices.size() * matrix_.log_rows());
        if (out.empty()) return out;                      // nothing to open (and out.data() may be null)
        matrix_.context().check(ss_merkle_open(matrix_.context().get(), tree_.get(), indices.data(), indices.size(), out.data()->data()));
        return out;
    }
    // MatrixMerkleTree::prove_rows(indices): (rows, paths)
    std::pair<std::vector<Felt>, std::vector<Digest>> prove_rows(const std::vector<uint64_t> &indices) const {
        return {matrix_.rows(indices), prove(indices)};
    }

  private:
    MatrixMerkleTree(const Matrix &m, ss_tree *t) : matrix_(m), tree_(t, [](ss_tree *p) { ss_tree_free(p); }) {}
    Matrix matrix_;                                   // keeps the committed columns alive (trees outlive prove(), §8b)
    std::shared_ptr<ss_tree> tree_;
};

}  // namespace sandstorm_b200
