// Pedersen hash over the StarkWare curve y^2 = x^3 + x + beta on Fp252 — the "algebraic" node hash of
// FriendlyMerkleTree's top layers (reference builtins/src/pedersen/mod.rs:31-36 -> starknet-crypto
// 0.6.1 pedersen_hash; points P0..P4 builtins/src/pedersen/constants.rs:6-29; curve
// builtins/src/utils.rs:141-152).
//
//   H(a,b) = [ P0 + a_lo*P1 + a_hi*P2 + b_lo*P3 + b_hi*P4 ]_x ,  lo = low 248 bits, hi = top 4 bits
//
// Device algorithm: fixed-base windowed sum with 8-bit windows.  The table holds d * 2^(8w) * P for
// every window w and digit d in 1..255 as affine points (1 MiB, L2-resident); one hash = at most 64
// Jacobian+affine additions (11 mults each) and one Fermat inversion done with a 250-squaring chain.
#pragma once
#include "fp252.cuh"
#include <vector>

namespace ss {

struct AffinePt { Fp x, y; };
struct JacPt { Fp x, y, z; };          // lazy-domain coordinates; z == 0 (mod p) <=> infinity

namespace ec {

constexpr int PED_WINDOW_BITS = 8;
constexpr int PED_LOW_WINDOWS = 31;                     // 31 * 8 = 248 low bits
constexpr int PED_DIGITS = (1 << PED_WINDOW_BITS) - 1;  // 255 non-zero digits
// table layout per input e in {0,1}: [31 windows][255 digits] for the low part, then [15] for the 4 high bits
constexpr int PED_POINTS_PER_INPUT = PED_LOW_WINDOWS * PED_DIGITS + 15;
constexpr int PED_TABLE_POINTS = 2 * PED_POINTS_PER_INPUT;

SS_HD bool is_zero_mod_p(const Fp &a) { return fp::is_zero_canon(fp::canon(a)); }

SS_HD JacPt jac_double(const JacPt &p) {
    // a = 1:  M = 3 X^2 + Z^4 ; S = 4 X Y^2 ; X3 = M^2 - 2S ; Y3 = M (S - X3) - 8 Y^4 ; Z3 = 2 Y Z
    const Fp xx = fp::sqr(p.x), yy = fp::sqr(p.y), yyyy = fp::sqr(yy), zz = fp::sqr(p.z);
    Fp s = fp::mul(p.x, yy);
    s = fp::add(s, s); s = fp::add(s, s);
    Fp m = fp::add(fp::add(xx, xx), xx);
    m = fp::add(m, fp::sqr(zz));
    JacPt r;
    r.x = fp::sub(fp::sub(fp::sqr(m), s), s);
    Fp y8 = fp::add(yyyy, yyyy); y8 = fp::add(y8, y8); y8 = fp::add(y8, y8);
    r.y = fp::sub(fp::mul(m, fp::sub(s, r.x)), y8);
    const Fp yz = fp::mul(p.y, p.z);
    r.z = fp::add(yz, yz);
    return r;
}

// p + q, q affine and not infinity.  Handles p == infinity, p == q (doubling), p == -q (infinity).
SS_HD JacPt jac_add_affine(const JacPt &p, const AffinePt &q) {
    if (is_zero_mod_p(p.z)) { JacPt r; r.x = q.x; r.y = q.y; r.z = fp::one(); return r; }
    const Fp z1z1 = fp::sqr(p.z);
    const Fp u2 = fp::mul(q.x, z1z1);
    const Fp s2 = fp::mul(fp::mul(q.y, p.z), z1z1);
    const Fp h = fp::sub(u2, p.x);
    const Fp rr = fp::sub(s2, p.y);
    if (is_zero_mod_p(h)) {
        if (is_zero_mod_p(rr)) return jac_double(p);
        JacPt inf; inf.x = fp::one(); inf.y = fp::one(); inf.z = fp::zero(); return inf;
    }
    const Fp hh = fp::sqr(h), hhh = fp::mul(hh, h), v = fp::mul(p.x, hh);
    JacPt r;
    r.x = fp::sub(fp::sub(fp::sub(fp::sqr(rr), hhh), v), v);
    r.y = fp::sub(fp::mul(rr, fp::sub(v, r.x)), fp::mul(p.y, hhh));
    r.z = fp::mul(p.z, h);
    return r;
}

// a^(p-2) with p - 2 = (2^59 + 2^4) * 2^192 + (2^192 - 1): 250 squarings + 11 multiplications
SS_HD Fp inv_chain(const Fp &x) {
    auto sqn = [](Fp v, int n) { for (int i = 0; i < n; ++i) v = fp::sqr(v); return v; };
    const Fp e2 = fp::mul(sqn(x, 1), x);
    const Fp e4 = fp::mul(sqn(e2, 2), e2);
    const Fp e8 = fp::mul(sqn(e4, 4), e4);
    const Fp e16 = fp::mul(sqn(e8, 8), e8);
    const Fp e32 = fp::mul(sqn(e16, 16), e16);
    const Fp e64 = fp::mul(sqn(e32, 32), e32);
    const Fp e128 = fp::mul(sqn(e64, 64), e64);
    const Fp e192 = fp::mul(sqn(e128, 64), e64);        // x^(2^192 - 1)
    const Fp g = fp::mul(e192, x);                       // x^(2^192)
    const Fp gh = sqn(fp::mul(sqn(g, 55), g), 4);        // g^(2^59 + 2^4)
    return fp::mul(gh, e192);
}

// digits of the canonical integer c (8 x u32 limbs): byte w of the low 248 bits, and the top nibble
SS_HD uint32_t ped_digit(const Fp &c, int w) { return (c.l[w >> 2] >> ((w & 3) * 8)) & 0xffu; }
SS_HD uint32_t ped_high(const Fp &c) { return (c.l[7] >> 24) & 0xfu; }

// The sum P0 + a_lo P1 + a_hi P2 + b_lo P3 + b_hi P4 in Jacobian coordinates: the hash is X / Z^2.  Callers that hash many
// pairs at once invert the Z's together (one inversion per thread block: Montgomery's trick, block_inverse below) instead
// of paying a 250-squaring Fermat chain per hash — a quarter of the multiplications of a hash.
template <typename LoadPt>
SS_HD JacPt pedersen_sum(const Fp &a_mont, const Fp &b_mont, const AffinePt &p0, LoadPt load_pt) {
    JacPt acc; acc.x = p0.x; acc.y = p0.y; acc.z = fp::one();
    Fp one_int = fp::zero(); one_int.l[0] = 1;
#pragma unroll 1
    for (int e = 0; e < 2; ++e) {
        const Fp c = fp::canon(fp::mul(e == 0 ? a_mont : b_mont, one_int));
        const int base = e * PED_POINTS_PER_INPUT;
#pragma unroll 1
        for (int w = 0; w < PED_LOW_WINDOWS; ++w) {
            const uint32_t d = ped_digit(c, w);
            if (d) acc = jac_add_affine(acc, load_pt(base + w * PED_DIGITS + (int)d - 1));
        }
        const uint32_t hi = ped_high(c);
        if (hi) acc = jac_add_affine(acc, load_pt(base + PED_LOW_WINDOWS * PED_DIGITS + (int)hi - 1));
    }
    return acc;
}

#ifdef __CUDACC__
// Inverse of every thread's value with ONE field inversion per block: up-sweep of pairwise products in shared memory,
// thread 0 inverts the root, down-sweep.  Every thread of the block must call it; sm holds 2 * THREADS elements.
template <int THREADS>
__device__ __forceinline__ Fp block_inverse(const Fp &v, Fp *sm) {
    const int tid = threadIdx.x;
    sm[THREADS + tid] = v;
    __syncthreads();
    for (int w = THREADS / 2; w >= 1; w >>= 1) {
        if (tid < w) sm[w + tid] = fp::mul(sm[2 * (w + tid)], sm[2 * (w + tid) + 1]);
        __syncthreads();
    }
    if (tid == 0) sm[1] = inv_chain(sm[1]);
    __syncthreads();
    for (int w = 1; w < THREADS; w <<= 1) {
        if (tid < w) {
            const int k = w + tid;
            const Fp ik = sm[k], l = sm[2 * k], r = sm[2 * k + 1];
            sm[2 * k] = fp::mul(ik, r);
            sm[2 * k + 1] = fp::mul(ik, l);
        }
        __syncthreads();
    }
    const Fp out = sm[THREADS + tid];
    __syncthreads();                                   // sm may be reused by the next call
    return out;
}
#endif

// Montgomery-form inputs/outputs.  `table` as laid out above, `p0` = shift point.
template <typename LoadPt>
SS_HD Fp pedersen_hash(const Fp &a_mont, const Fp &b_mont, const AffinePt &p0, LoadPt load_pt) {
    JacPt acc; acc.x = p0.x; acc.y = p0.y; acc.z = fp::one();
    Fp one_int = fp::zero(); one_int.l[0] = 1;
#pragma unroll 1
    for (int e = 0; e < 2; ++e) {
        // canonical integer = mont * R^-1 : Montgomery-multiply by the plain integer 1
        const Fp c = fp::canon(fp::mul(e == 0 ? a_mont : b_mont, one_int));
        const int base = e * PED_POINTS_PER_INPUT;
#pragma unroll 1
        for (int w = 0; w < PED_LOW_WINDOWS; ++w) {
            const uint32_t d = ped_digit(c, w);
            if (d) acc = jac_add_affine(acc, load_pt(base + w * PED_DIGITS + (int)d - 1));
        }
        const uint32_t hi = ped_high(c);
        if (hi) acc = jac_add_affine(acc, load_pt(base + PED_LOW_WINDOWS * PED_DIGITS + (int)hi - 1));
    }
    const Fp zi = inv_chain(acc.z);
    return fp::canon(fp::mul(acc.x, fp::sqr(zi)));
}

// ----------------------------------------------------------------- host-side table generation
const uint32_t PED_BASE[5][2][8] = {
    // P0..P4 (canonical integers, u32 little-endian limbs): builtins/src/pedersen/constants.rs:6-29
    {{0x50ca6804u, 0x551fde40u, 0x22947733u, 0x716b0b10u, 0xeb599f16u, 0x00ee1b87u, 0xa8c16007u, 0x049ee3ebu},
     {0x6e10268au, 0xd0405d26u, 0xc0e056c1u, 0x4e621062u, 0x06ea0ed3u, 0xf346d49du, 0x4b3bc6ddu, 0x03ca0cfeu}},
    {{0x57ebe47bu, 0x1080d179u, 0x6d56eb0cu, 0x8fa8120bu, 0x55fca9e5u, 0x969c7486u, 0xcbaffe7fu, 0x0234287du},
     {0xe89e5615u, 0x6ed0268eu, 0x7a6c94ccu, 0x940135ddu, 0xd41f4e39u, 0x1e889527u, 0x00f96fb2u, 0x03b056f1u}},
    {{0xba8aa378u, 0xb7a6932du, 0xde5e3018u, 0x99099ec1u, 0x56558f33u, 0x3f9dab26u, 0x76c83db3u, 0x04fa56f3u},
     {0x0ff5b54du, 0x5168f4e8u, 0x2a7a23b4u, 0x562761f9u, 0xe47e4401u, 0x8113e0c0u, 0xc931c9e3u, 0x03fa0984u}},
    {{0xbd2d6997u, 0x3aa372f0u, 0x4709e90fu, 0x40c690c7u, 0x5b45f74bu, 0x764910f7u, 0x66be8decu, 0x04ba4cc1u},
     {0xb24b219cu, 0x48151f27u, 0x5ce5ae7cu, 0xcac5c59au, 0xc4ede85fu, 0x4b971e46u, 0xf5c1751fu, 0x0040301cu}},
    {{0x49a58202u, 0xd36ff12cu, 0xd53fb325u, 0x2ca65048u, 0xf61a63bbu, 0x6e44cca8u, 0xb0e6cc1cu, 0x054302dcu},
     {0xe99c2426u, 0x879dcc77u, 0x3c25561au, 0xce98ad78u, 0x68d8ae25u, 0xb3480462u, 0x37d13504u, 0x01b77b3eu}},
};

inline AffinePt base_point(int k) {
    AffinePt p;
    Fp cx, cy;
    for (int i = 0; i < 8; ++i) { cx.l[i] = PED_BASE[k][0][i]; cy.l[i] = PED_BASE[k][1][i]; }
    p.x = fp::canon(fp::mul(cx, fp::r2()));
    p.y = fp::canon(fp::mul(cy, fp::r2()));
    return p;
}

inline AffinePt to_affine(const JacPt &p) {
    const Fp zi = inv_chain(p.z), zi2 = fp::sqr(zi);
    AffinePt a;
    a.x = fp::canon(fp::mul(p.x, zi2));
    a.y = fp::canon(fp::mul(p.y, fp::mul(zi2, zi)));
    return a;
}

// fills PED_TABLE_POINTS + 1 affine points (2 Fp each); signature matches cached_table's callback
inline void fill_pedersen_table(Fp *dst, size_t n_fp, int, int) {
    const size_t n_pts = n_fp / 2;
    std::vector<JacPt> jac(n_pts);
    size_t k = 0;
    for (int e = 0; e < 2; ++e) {
        AffinePt bw = base_point(1 + 2 * e);                       // P1 / P3: low 248 bits
        for (int w = 0; w < PED_LOW_WINDOWS; ++w) {
            JacPt s; s.x = bw.x; s.y = bw.y; s.z = fp::one();
            for (int d = 1; d <= PED_DIGITS; ++d) {
                jac[k++] = s;
                s = jac_add_affine(s, bw);
            }
            JacPt nb; nb.x = bw.x; nb.y = bw.y; nb.z = fp::one();
            for (int t = 0; t < PED_WINDOW_BITS; ++t) nb = jac_double(nb);
            bw = to_affine(nb);
        }
        const AffinePt hi = base_point(2 + 2 * e);                  // P2 / P4: top 4 bits
        JacPt s; s.x = hi.x; s.y = hi.y; s.z = fp::one();
        for (int d = 1; d <= 15; ++d) {
            jac[k++] = s;
            s = jac_add_affine(s, hi);
        }
    }
    const AffinePt p0 = base_point(0);
    jac[k].x = p0.x; jac[k].y = p0.y; jac[k].z = fp::one();
    ++k;
    // batch conversion to affine (Montgomery's trick on the z coordinates)
    std::vector<Fp> prefix(n_pts);
    Fp acc = fp::one();
    for (size_t i = 0; i < n_pts; ++i) { prefix[i] = acc; acc = fp::mul(acc, jac[i].z); }
    Fp inv = inv_chain(acc);
    for (size_t i = n_pts; i-- > 0;) {
        const Fp zi = fp::mul(inv, prefix[i]);
        inv = fp::mul(inv, jac[i].z);
        const Fp zi2 = fp::sqr(zi);
        dst[2 * i] = fp::canon(fp::mul(jac[i].x, zi2));
        dst[2 * i + 1] = fp::canon(fp::mul(jac[i].y, fp::mul(zi2, zi)));
    }
}


}  // namespace ec
}  // namespace ss
