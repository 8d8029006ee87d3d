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
void hc_fp_from_u32(uint32_t v, uint32_t *out) {
    Fp r = fp::from_u32(v); std::memcpy(out, r.l, 32);
}

}  // extern "C"
