// Carry-chain primitives.  On the device each function is exactly one PTX instruction
// (ptxas fuses mad.lo.cc / madc.hi.cc pairs on the same operands into IMAD.WIDE.U32 with
// carry predicates).  On the host the same functions are emulated with a thread-local carry
// flag so that the field arithmetic built on them can be unit-tested in the CPU-only build
// container (tests/test_host_arith.py) before it is run on a B200.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define SS_HD __host__ __device__ __forceinline__
#define SS_D __device__ __forceinline__
#else
#define SS_HD inline
#define SS_D inline
#endif

namespace ss {
namespace ptx {

#if defined(__CUDA_ARCH__)

SS_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r; }
SS_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SS_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SS_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }
SS_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r; }

#else  // host emulation

inline uint32_t &cf() { static thread_local uint32_t f = 0; return f; }
inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b; cf() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t s = (uint64_t)a + b + cf(); cf() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cf(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b; cf() = (uint32_t)((d >> 32) & 1); return (uint32_t)d; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t d = (uint64_t)a - b - cf(); cf() = (uint32_t)((d >> 32) & 1); return (uint32_t)d; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cf(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c; cf() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_lo(a, b) + c + cf(); cf() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint64_t s = (uint64_t)mul_hi(a, b) + c + cf(); cf() = (uint32_t)(s >> 32); return (uint32_t)s; }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return mul_hi(a, b) + c + cf(); }

#endif


// acc[0..8] += (a0, a1, a2, a3) * b where product k lands on limbs (2k, 2k+1); acc[8] takes the carry.
// One asm block so that ptxas keeps each (lo,hi) pair in an aligned register pair and emits
// IMAD.WIDE.U32 / IMAD.WIDE.U32.X with carry predicates (4 instructions + 1 IADD3.X).
SS_HD void mad4_chain(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t &r5,
                      uint32_t &r6, uint32_t &r7, uint32_t &r8, uint32_t a0, uint32_t a1, uint32_t a2,
                      uint32_t a3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %9, %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9, %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32 %8, %8, 0;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7), "+r"(r8)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    r0 = mad_lo_cc(a0, b, r0); r1 = madc_hi_cc(a0, b, r1);
    r2 = madc_lo_cc(a1, b, r2); r3 = madc_hi_cc(a1, b, r3);
    r4 = madc_lo_cc(a2, b, r4); r5 = madc_hi_cc(a2, b, r5);
    r6 = madc_lo_cc(a3, b, r6); r7 = madc_hi_cc(a3, b, r7);
    r8 = addc(r8, 0);
#endif
}
// same without the carry limb (top row, where the carry out is provably zero)
SS_HD void mad4_chain_top(uint32_t &r0, uint32_t &r1, uint32_t &r2, uint32_t &r3, uint32_t &r4, uint32_t &r5,
                          uint32_t &r6, uint32_t &r7, uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                          uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %8, %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32 %7, %11, %12, %7;"
        : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4), "+r"(r5), "+r"(r6), "+r"(r7)
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b));
#else
    r0 = mad_lo_cc(a0, b, r0); r1 = madc_hi_cc(a0, b, r1);
    r2 = madc_lo_cc(a1, b, r2); r3 = madc_hi_cc(a1, b, r3);
    r4 = madc_lo_cc(a2, b, r4); r5 = madc_hi_cc(a2, b, r5);
    r6 = madc_lo_cc(a3, b, r6); r7 = madc_hi(a3, b, r7);
#endif
}

}  // namespace ptx
}  // namespace ss
