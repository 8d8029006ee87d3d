// CPU build of the __host__ __device__ arithmetic (PTX carry primitives emulated, see arith.cuh)
// so that tests can exercise the exact device algorithms in the GPU-less build container.
// This is a TEST HOOK: it exports no ss_* product entry point and is never used by the product path.
#include "fp252.cuh"
#include <cstring>

using namespace ss;

extern "C" {

void hc_fp_mul(const uint32_t *a, const uint32_t *b, uint32_t *out, int canonical) {
    Fp x, y;
    std::memcpy(x.l, a, 32);
    std::memcpy(y.l, b, 32);
    Fp r = fp::mul(x, y);
    if (canonical) r = fp::canon(r);
    std::memcpy(out, r.l, 32);
}
void hc_fp_add(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    Fp x, y; std::memcpy(x.l, a, 32); std::memcpy(y.l, b, 32);
    Fp r = fp::add(x, y); std::memcpy(out, r.l, 32);
}
void hc_fp_sub(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    Fp x, y; std::memcpy(x.l, a, 32); std::memcpy(y.l, b, 32);
    Fp r = fp::sub(x, y); std::memcpy(out, r.l, 32);
}
void hc_fp_sub4p(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    Fp x, y; std::memcpy(x.l, a, 32); std::memcpy(y.l, b, 32);
    Fp r = fp::sub4p(x, y); std::memcpy(out, r.l, 32);
}
void hc_fp_sub2p(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    Fp x, y; std::memcpy(x.l, a, 32); std::memcpy(y.l, b, 32);
    Fp r = fp::sub2p(x, y); std::memcpy(out, r.l, 32);
}
void hc_fp_canon(const uint32_t *a, uint32_t *out) {
    Fp x; std::memcpy(x.l, a, 32);
    Fp r = fp::canon(x); std::memcpy(out, r.l, 32);
}
void hc_fp_reduce8p(const uint32_t *a, uint32_t *out) {
    Fp x; std::memcpy(x.l, a, 32);
    fp::cond_sub_4p(x); fp::cond_sub_2p(x); std::memcpy(out, x.l, 32);
}
void hc_fp_inv(const uint32_t *a, uint32_t *out) {
    Fp x; std::memcpy(x.l, a, 32);
    Fp r = fp::canon(fp::inv(x)); std::memcpy(out, r.l, 32);
}
void hc_fp_red(const uint32_t *a, uint32_t *out) {
    Fp x; std::memcpy(x.l, a, 32);
    Fp r = fp::red(x); std::memcpy(out, r.l, 32);
}
void hc_fp_sub_kp(const uint32_t *a, const uint32_t *b, uint32_t k, uint32_t *out) {
    Fp x, y; std::memcpy(x.l, a, 32); std::memcpy(y.l, b, 32);
    Fp r = fp::sub_kp(x, y, k); std::memcpy(out, r.l, 32);
}
// out = Montgomery reduction of sum_k a[k] * b[k] (n terms of 8 limbs each)
void hc_fp_dot(const uint32_t *a, const uint32_t *b, int n, uint32_t *out) {
    fp::WideAcc w;
    fp::acc_init(w);
    for (int k = 0; k < n; ++k) {
        Fp x, y; std::memcpy(x.l, a + 8 * k, 32); std::memcpy(y.l, b + 8 * k, 32);
        fp::acc_mac(w, x, y);
    }
    Fp r = fp::acc_reduce(w); std::memcpy(out, r.l, 32);
}
void hc_fp_from_u32(uint32_t v, uint32_t *out) {
    Fp r = fp::from_u32(v); std::memcpy(out, r.l, 32);
}

}  // extern "C"

// ---- hash cores and Pedersen (host build of the device code) ----
#include "hashes.cuh"
#include "pedersen.cuh"
#include <vector>

extern "C" {

// msg: n_bytes (multiple of 8) -> Keccak-256 digest
void hc_keccak256(const uint8_t *msg, int n_bytes, uint8_t *out) {
    const uint64_t *l = reinterpret_cast<const uint64_t *>(msg);
    auto lane = [&](int g) -> uint64_t { uint64_t v; std::memcpy(&v, msg + 8 * g, 8); (void)l; return v; };
    uint64_t d[4];
    hash::keccak256_lanes(lane, n_bytes / 8, d);
    std::memcpy(out, d, 32);
}
void hc_blake2s256(const uint8_t *msg, int n_bytes, uint8_t *out) {
    auto word = [&](int g) -> uint32_t { uint32_t v; std::memcpy(&v, msg + 4 * g, 4); return v; };
    uint32_t h[8];
    hash::blake2s256_words(word, n_bytes / 4, h);
    std::memcpy(out, h, 32);
}
void hc_sha256(const uint8_t *msg, int n_bytes, uint8_t *out) {
    auto word = [&](int g) -> uint32_t { return ((uint32_t)msg[4 * g] << 24) | ((uint32_t)msg[4 * g + 1] << 16) | ((uint32_t)msg[4 * g + 2] << 8) | msg[4 * g + 3]; };
    uint32_t h[8];
    hash::sha256_words(word, n_bytes / 4, h);
    for (int i = 0; i < 8; ++i) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
void hc_pedersen(const uint32_t *a, const uint32_t *b, uint32_t *out) {
    static std::vector<Fp> table;
    if (table.empty()) {
        table.resize(2 * (size_t)(ec::PED_TABLE_POINTS + 1));
        ec::fill_pedersen_table(table.data(), table.size(), 0, 0);
    }
    const AffinePt *pts = reinterpret_cast<const AffinePt *>(table.data());
    Fp x, y;
    std::memcpy(x.l, a, 32);
    std::memcpy(y.l, b, 32);
    const Fp h = ec::pedersen_hash(x, y, pts[ec::PED_TABLE_POINTS], [&](int i) { return pts[i]; });
    std::memcpy(out, h.l, 32);
}
void hc_inv_chain(const uint32_t *a, uint32_t *out) {
    Fp x; std::memcpy(x.l, a, 32);
    Fp r = fp::canon(ec::inv_chain(x)); std::memcpy(out, r.l, 32);
}

}  // extern "C"

// ---- Goldilocks single-limb arithmetic (goldilocks.cuh) ----
#include "goldilocks.cuh"
extern "C" {
uint64_t hc_gl_add(uint64_t a, uint64_t b) { return gl::add(a, b); }
uint64_t hc_gl_sub(uint64_t a, uint64_t b) { return gl::sub(a, b); }
uint64_t hc_gl_mul(uint64_t a, uint64_t b) { return gl::mul(a, b); }
uint64_t hc_gl_reduce128(uint64_t lo, uint64_t hi) { return gl::reduce128(lo, hi); }
uint64_t hc_gl_root(int log_n) { return gl::root_of_unity(log_n); }
}
