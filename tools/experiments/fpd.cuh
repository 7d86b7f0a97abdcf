// fpd.cuh -- the same Montgomery arithmetic in Fp as fp.cuh, carried by the FP64 pipe instead of the integer-multiply pipe.
//
// Why: the dominant kernel (k_g2_decompress, two 380-bit Fp exponentiations per signature) runs the `fmaheavy` pipe --
// IMAD.WIDE.U32, one warp instruction per 4 cycles per scheduler -- at ~90 % of its peak (profiles/README.md), while the FP64
// pipe of the same SM (DFMA, 64 lanes/clk/SM on B200) idles.  The two pipes issue independently, so warps that run the
// exponentiation chain in double precision ADD throughput instead of competing for it.  Only the chain needs this form:
// a chain is multiplications and squarings only, entered and left through fp <-> fpd conversions.
//
// Representation: 16 digits of 24 bits held as doubles (exact integers), value = sum l[i] * 2^(24 i), Montgomery form with
// the SAME R = 2^384 as fp.cuh (16 * 24 = 12 * 32), so both forms of a chain produce identical canonical limbs.  Digits are
// SIGNED, |l[i]| <= 2^23 + 2^6 after fpd_normalize (l[15] carries the sign of the value), and values are only bounded,
// |v| < 0.63 p, not canonical: Montgomery reduction is indifferent to the sign of its input.
//
// Exactness: every intermediate is an integer of magnitude < 2^53.  A column of the product collects <= 16 digit products
// (<= 16 * (2^23+2^6)^2 < 2^50.1) plus <= 16 products m_i * p_j with |m_i| <= 2^23, p_j < 2^24 (<= 2^51) plus one carry
// (< 2^29): < 2^52.  Rounding x/2^24 to an integer uses the 1.5*2^52 magic constant (|x| 2^-24 < 2^51 holds with room).
#pragma once
#include "fp.cuh"
#ifndef __CUDA_ARCH__
#include <cmath>
#endif

namespace b2 {

struct fpd {
    double l[16];
};

#define FPD_MAGIC 6755399441055744.0          /* 1.5 * 2^52 */
#define FPD_2P24 16777216.0                   /* 2^24  */
#define FPD_2M24 5.9604644775390625e-08       /* 2^-24 */
#define FPD_PINV (-196611.0)                  /* -p^-1 mod 2^24 = 0xfcfffd, taken in (-2^23, 2^23) */

HD double d_fma(double a, double b, double c) {
#ifdef __CUDA_ARCH__
    return __fma_rn(a, b, c);
#else
    return std::fma(a, b, c);
#endif
}
HD double d_add(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dadd_rn(a, b);                  // never contracted into an fma
#else
    volatile double r = a + b;
    return r;
#endif
}
HD double d_mul(double a, double b) {
#ifdef __CUDA_ARCH__
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
// x = q * 2^24 + r,  q = round-to-nearest(x / 2^24),  |r| <= 2^23
HD void d_split(double x, double& q, double& r) {
    q = d_add(d_fma(x, FPD_2M24, FPD_MAGIC), -FPD_MAGIC);
    r = d_fma(q, -FPD_2P24, x);
}

// 24-bit digit j of p
HD constexpr uint32_t P24U(int j) {
    return (uint32_t)((((uint64_t)P_LIMB((24 * j) / 32) | ((uint64_t)(((24 * j) / 32 + 1 < 12) ? P_LIMB(((24 * j) / 32 + 1) % 12) : 0u) << 32)) >>
                       ((24 * j) % 32)) &
                      0xffffffu);
}
HD constexpr double P24D(int j) { return (double)P24U(j); }

// carries: four independent 4-digit chains, then the three chain-to-chain carries.  In: |l[i]| < 2^52.  Out: |l[i]| <= 2^23+2^6
// for i < 15; l[15] keeps whatever is left (small, because |v| < 2^381).
HD void fpd_normalize(double* t) {
    double q, c3, c7, c11;
#pragma unroll
    for (int s = 0; s < 4; s++) {
#pragma unroll
        for (int j = 4 * s; j < 4 * s + 3; j++) {
            d_split(t[j], q, t[j]);
            t[j + 1] = d_add(t[j + 1], q);
        }
    }
    d_split(t[3], c3, t[3]);
    d_split(t[7], c7, t[7]);
    d_split(t[11], c11, t[11]);
    t[4] = d_add(t[4], c3);
    t[8] = d_add(t[8], c7);
    t[12] = d_add(t[12], c11);
    d_split(t[4], q, t[4]);
    t[5] = d_add(t[5], q);
    d_split(t[8], q, t[8]);
    t[9] = d_add(t[9], q);
    d_split(t[12], q, t[12]);
    t[13] = d_add(t[13], q);
}

// Montgomery reduction of a sliding window of 16 columns.  COL(k) yields column k of the double-length product (k = 0..30).
// Row i: m = t[0] * (-p^-1) mod 2^24 (signed), t += m * p, t[0] is now a multiple of 2^24 and moves into t[1] as a carry; the window
// slides by one digit and column i+16 of the product enters at the top.
template <class COL>
HD fpd fpd_mont(COL col) {
    double t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = col(k);
#pragma unroll
    for (int i = 0; i < 16; i++) {
        double q, lo, m;
        d_split(t[0], q, lo);
        d_split(d_mul(lo, FPD_PINV), q, m);
#pragma unroll
        for (int j = 0; j < 16; j++) t[j] = d_fma(m, P24D(j), t[j]);
        t[1] = d_fma(t[0], FPD_2M24, t[1]);
#pragma unroll
        for (int j = 0; j < 15; j++) t[j] = t[j + 1];
        t[15] = (i + 16 <= 30) ? col(i + 16) : 0.0;
    }
    fpd_normalize(t);
    fpd r;
#pragma unroll
    for (int k = 0; k < 16; k++) r.l[k] = t[k];
    return r;
}

struct fpd_mul_col {
    const fpd& a;
    const fpd& b;
    HD double operator()(int k) const {
        const int i0 = k < 16 ? 0 : k - 15, i1 = k < 16 ? k : 15;
        double s = d_mul(a.l[i0], b.l[k - i0]);
#pragma unroll
        for (int i = i0 + 1; i <= i1; i++) s = d_fma(a.l[i], b.l[k - i], s);
        return s;
    }
};
struct fpd_sqr_col {
    const fpd& a;
    HD double operator()(int k) const {
        const int i0 = k < 16 ? 0 : k - 15;          // pairs (i, k-i) with i < k-i
        double s = 0.0;
        bool first = true;
#pragma unroll
        for (int i = i0; 2 * i < k; i++) {
            s = first ? d_mul(a.l[i], a.l[k - i]) : d_fma(a.l[i], a.l[k - i], s);
            first = false;
        }
        s = d_add(s, s);
        if ((k & 1) == 0) s = d_fma(a.l[k / 2], a.l[k / 2], s);
        return s;
    }
};

HD fpd fpd_mul(const fpd& a, const fpd& b) { return fpd_mont(fpd_mul_col{a, b}); }
HD fpd fpd_sqr(const fpd& a) { return fpd_mont(fpd_sqr_col{a}); }

// ---- conversions (canonical Montgomery limbs <-> signed digits)
HD fpd fpd_from_fp(const fp& a) {
    uint32_t d[16];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const uint32_t w0 = a.l[3 * k], w1 = a.l[3 * k + 1], w2 = a.l[3 * k + 2];
        d[4 * k] = w0 & 0xffffffu;
        d[4 * k + 1] = ((w0 >> 24) | (w1 << 8)) & 0xffffffu;
        d[4 * k + 2] = ((w1 >> 16) | (w2 << 16)) & 0xffffffu;
        d[4 * k + 3] = w2 >> 8;
    }
    double t[16];
#pragma unroll
    for (int k = 0; k < 16; k++) t[k] = (double)(int32_t)d[k];
    fpd_normalize(t);
    fpd r;
#pragma unroll
    for (int k = 0; k < 16; k++) r.l[k] = t[k];
    return r;
}
HD fp fpd_to_fp(const fpd& a) {
    // v + p is in (0, 2p): propagate carries in integers, pack, subtract p once if needed
    uint32_t d[16];
    int32_t carry = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) {
#ifdef __CUDA_ARCH__
        int32_t v = __double2int_rn(a.l[k]);
#else
        int32_t v = (int32_t)a.l[k];
#endif
        v += (int32_t)P24U(k) + carry;
        if (k < 15) {
            d[k] = (uint32_t)v & 0xffffffu;
            carry = v >> 24;
        } else {
            d[k] = (uint32_t)v;
        }
    }
    fp r;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        r.l[3 * k] = d[4 * k] | (d[4 * k + 1] << 24);
        r.l[3 * k + 1] = (d[4 * k + 1] >> 8) | (d[4 * k + 2] << 16);
        r.l[3 * k + 2] = (d[4 * k + 2] >> 16) | (d[4 * k + 3] << 8);
    }
    fp_final_sub(r.l);
    return r;
}

// the sliding-window exponentiation of fp_pow_prog, FP64 form; same program table, bit-identical result
HDN fp fpd_pow_prog(const fp& a_in, int off) {
    const uint32_t* prog = const_table() + off;
    fpd tbl[8];
    const fpd a = fpd_from_fp(a_in);
    const fpd a2 = fpd_sqr(a);
    tbl[0] = a;
#pragma unroll 1
    for (int i = 1; i < 8; i++) tbl[i] = fpd_mul(tbl[i - 1], a2);
    const uint32_t n = prog[0];
    fpd r = tbl[prog[1] & 0xffu];
#pragma unroll 1
    for (uint32_t k = 2; k <= n; k++) {
        const uint32_t op = prog[k];
#pragma unroll 1
        for (uint32_t s = op >> 8; s; s--) r = fpd_sqr(r);
        const uint32_t idx = op & 0xffu;
        if (idx != 0xffu) r = fpd_mul(r, tbl[idx]);
    }
    return fpd_to_fp(r);
}

}  // namespace b2
