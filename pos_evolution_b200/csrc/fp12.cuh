// fp12.cuh -- the degree-6 and degree-12 extensions used by the pairing:
//   Fp6  = Fp2[v]/(v^3 - xi), xi = 1+i      element a0 + a1 v + a2 v^2
//   Fp12 = Fp6[w]/(w^2 - v)                 element b0 + b1 w
// Same tower as the oracle (oracle/bls12_381.py) so that pairing values can be compared
// coefficient by coefficient.  Part of kernel library K1.
#pragma once
#include "fp2.cuh"

namespace b2 {

struct fp6 {
    fp2 a0, a1, a2;
};
struct fp12 {
    fp6 b0, b1;
};

HD fp6 fp6_zero() {
    fp6 r;
    r.a0 = fp2_zero();
    r.a1 = fp2_zero();
    r.a2 = fp2_zero();
    return r;
}
HD fp6 fp6_one() {
    fp6 r = fp6_zero();
    r.a0 = fp2_one();
    return r;
}
HD fp6 fp6_add(const fp6& x, const fp6& y) {
    fp6 r;
    r.a0 = fp2_add(x.a0, y.a0);
    r.a1 = fp2_add(x.a1, y.a1);
    r.a2 = fp2_add(x.a2, y.a2);
    return r;
}
HD fp6 fp6_sub(const fp6& x, const fp6& y) {
    fp6 r;
    r.a0 = fp2_sub(x.a0, y.a0);
    r.a1 = fp2_sub(x.a1, y.a1);
    r.a2 = fp2_sub(x.a2, y.a2);
    return r;
}
HD fp6 fp6_neg(const fp6& x) {
    fp6 r;
    r.a0 = fp2_neg(x.a0);
    r.a1 = fp2_neg(x.a1);
    r.a2 = fp2_neg(x.a2);
    return r;
}
HD fp6 fp6_dbl(const fp6& x) { return fp6_add(x, x); }
HD bool fp6_eq(const fp6& x, const fp6& y) { return fp2_eq(x.a0, y.a0) && fp2_eq(x.a1, y.a1) && fp2_eq(x.a2, y.a2); }
// multiply by v: (a0, a1, a2) -> (xi*a2, a0, a1)
HD fp6 fp6_mul_v(const fp6& x) {
    fp6 r;
    r.a0 = fp2_mul_xi(x.a2);
    r.a1 = x.a0;
    r.a2 = x.a1;
    return r;
}
// Karatsuba, 6 Fp2 multiplications
HDN fp6 fp6_mul(const fp6& x, const fp6& y) {
    fp2 v0 = fp2_mul(x.a0, y.a0);
    fp2 v1 = fp2_mul(x.a1, y.a1);
    fp2 v2 = fp2_mul(x.a2, y.a2);
    fp2 t12 = fp2_mul(fp2_add(x.a1, x.a2), fp2_add(y.a1, y.a2));
    fp2 t01 = fp2_mul(fp2_add(x.a0, x.a1), fp2_add(y.a0, y.a1));
    fp2 t02 = fp2_mul(fp2_add(x.a0, x.a2), fp2_add(y.a0, y.a2));
    fp6 r;
    r.a0 = fp2_add(v0, fp2_mul_xi(fp2_sub(fp2_sub(t12, v1), v2)));
    r.a1 = fp2_add(fp2_sub(fp2_sub(t01, v0), v1), fp2_mul_xi(v2));
    r.a2 = fp2_add(fp2_sub(fp2_sub(t02, v0), v2), v1);
    return r;
}
// Chung-Hasan SQR2: 2 mul + 3 sqr in Fp2
HDN fp6 fp6_sqr(const fp6& x) {
    fp2 s0 = fp2_sqr(x.a0);
    fp2 s1 = fp2_dbl(fp2_mul(x.a0, x.a1));
    fp2 s2 = fp2_sqr(fp2_add(fp2_sub(x.a0, x.a1), x.a2));
    fp2 s3 = fp2_dbl(fp2_mul(x.a1, x.a2));
    fp2 s4 = fp2_sqr(x.a2);
    fp6 r;
    r.a0 = fp2_add(s0, fp2_mul_xi(s3));
    r.a1 = fp2_add(s1, fp2_mul_xi(s4));
    r.a2 = fp2_sub(fp2_sub(fp2_add(fp2_add(s1, s2), s3), s0), s4);
    return r;
}
// x * (b0 + b1 v): 5 Fp2 multiplications
HDN fp6 fp6_mul_by_01(const fp6& x, const fp2& b0, const fp2& b1) {
    fp2 v0 = fp2_mul(x.a0, b0);
    fp2 v1 = fp2_mul(x.a1, b1);
    fp2 t0 = fp2_mul(fp2_add(x.a1, x.a2), b1);                 // a1 b1 + a2 b1
    fp2 t1 = fp2_mul(fp2_add(x.a0, x.a1), fp2_add(b0, b1));     // a0b0 + a0b1 + a1b0 + a1b1
    fp2 t2 = fp2_mul(fp2_add(x.a0, x.a2), b0);                 // a0 b0 + a2 b0
    fp6 r;
    r.a0 = fp2_add(v0, fp2_mul_xi(fp2_sub(t0, v1)));
    r.a1 = fp2_sub(fp2_sub(t1, v0), v1);
    r.a2 = fp2_add(fp2_sub(t2, v0), v1);
    return r;
}
// x * (b1 v): 3 Fp2 multiplications
HDN fp6 fp6_mul_by_1(const fp6& x, const fp2& b1) {
    fp6 r;
    r.a0 = fp2_mul_xi(fp2_mul(x.a2, b1));
    r.a1 = fp2_mul(x.a0, b1);
    r.a2 = fp2_mul(x.a1, b1);
    return r;
}
HDN fp6 fp6_inv(const fp6& x) {
    fp2 t0 = fp2_sub(fp2_sqr(x.a0), fp2_mul_xi(fp2_mul(x.a1, x.a2)));
    fp2 t1 = fp2_sub(fp2_mul_xi(fp2_sqr(x.a2)), fp2_mul(x.a0, x.a1));
    fp2 t2 = fp2_sub(fp2_sqr(x.a1), fp2_mul(x.a0, x.a2));
    fp2 d = fp2_add(fp2_mul(x.a0, t0), fp2_mul_xi(fp2_add(fp2_mul(x.a2, t1), fp2_mul(x.a1, t2))));
    fp2 di = fp2_inv(d);
    fp6 r;
    r.a0 = fp2_mul(t0, di);
    r.a1 = fp2_mul(t1, di);
    r.a2 = fp2_mul(t2, di);
    return r;
}

// ------------------------------------------------------------------------------------------ Fp12
HD fp12 fp12_one() {
    fp12 r;
    r.b0 = fp6_one();
    r.b1 = fp6_zero();
    return r;
}
HD bool fp12_eq(const fp12& x, const fp12& y) { return fp6_eq(x.b0, y.b0) && fp6_eq(x.b1, y.b1); }
HD bool fp12_is_one(const fp12& x) { return fp12_eq(x, fp12_one()); }
HD fp12 fp12_conj(const fp12& x) {
    fp12 r;
    r.b0 = x.b0;
    r.b1 = fp6_neg(x.b1);
    return r;
}
HDN fp12 fp12_mul(const fp12& x, const fp12& y) {
    fp6 t0 = fp6_mul(x.b0, y.b0);
    fp6 t1 = fp6_mul(x.b1, y.b1);
    fp6 t2 = fp6_mul(fp6_add(x.b0, x.b1), fp6_add(y.b0, y.b1));
    fp12 r;
    r.b0 = fp6_add(t0, fp6_mul_v(t1));
    r.b1 = fp6_sub(fp6_sub(t2, t0), t1);
    return r;
}
// complex squaring: 2 Fp6 multiplications
HDN fp12 fp12_sqr(const fp12& x) {
    fp6 t = fp6_mul(x.b0, x.b1);
    fp6 s = fp6_mul(fp6_add(x.b0, x.b1), fp6_add(x.b0, fp6_mul_v(x.b1)));
    fp12 r;
    r.b0 = fp6_sub(fp6_sub(s, t), fp6_mul_v(t));
    r.b1 = fp6_dbl(t);
    return r;
}
// multiply by a sparse line value  (l0 + l1 v) + (l4 v) w   -- 13 Fp2 multiplications
HDN fp12 fp12_mul_by_014(const fp12& x, const fp2& l0, const fp2& l1, const fp2& l4) {
    fp6 t0 = fp6_mul_by_01(x.b0, l0, l1);
    fp6 t1 = fp6_mul_by_1(x.b1, l4);
    fp6 t2 = fp6_mul_by_01(fp6_add(x.b0, x.b1), l0, fp2_add(l1, l4));
    fp12 r;
    r.b0 = fp6_add(t0, fp6_mul_v(t1));
    r.b1 = fp6_sub(fp6_sub(t2, t0), t1);
    return r;
}
HDN fp12 fp12_inv(const fp12& x) {
    fp6 d = fp6_sub(fp6_sqr(x.b0), fp6_mul_v(fp6_sqr(x.b1)));
    fp6 di = fp6_inv(d);
    fp12 r;
    r.b0 = fp6_mul(x.b0, di);
    r.b1 = fp6_neg(fp6_mul(x.b1, di));
    return r;
}

// p-power Frobenius.  In the basis w^k (k = 0..5: a0, b1.a0, a1, b1.a1, a2, b1.a2 of the tower)
// coefficient g_k maps to conj(g_k) * gamma^k, gamma = xi^((p-1)/6)  (constants C_FROB1_k).
HDN fp12 fp12_frob(const fp12& x) {
    fp12 r;
    r.b0.a0 = fp2_conj(x.b0.a0);
    r.b1.a0 = fp2_mul(fp2_conj(x.b1.a0), fp2_load_const(C_FROB1_1));
    r.b0.a1 = fp2_mul(fp2_conj(x.b0.a1), fp2_load_const(C_FROB1_2));
    r.b1.a1 = fp2_mul(fp2_conj(x.b1.a1), fp2_load_const(C_FROB1_3));
    r.b0.a2 = fp2_mul(fp2_conj(x.b0.a2), fp2_load_const(C_FROB1_4));
    r.b1.a2 = fp2_mul(fp2_conj(x.b1.a2), fp2_load_const(C_FROB1_5));
    return r;
}
// p^2-power Frobenius: g_k -> g_k * norm(gamma^k) (constants C_FROB2_k, in Fp)
HDN fp12 fp12_frob2(const fp12& x) {
    fp12 r;
    r.b0.a0 = x.b0.a0;
    r.b1.a0 = fp2_mul_fp(x.b1.a0, fp_load_const(C_FROB2_1));
    r.b0.a1 = fp2_mul_fp(x.b0.a1, fp_load_const(C_FROB2_2));
    r.b1.a1 = fp2_mul_fp(x.b1.a1, fp_load_const(C_FROB2_3));
    r.b0.a2 = fp2_mul_fp(x.b0.a2, fp_load_const(C_FROB2_4));
    r.b1.a2 = fp2_mul_fp(x.b1.a2, fp_load_const(C_FROB2_5));
    return r;
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the
// final exponentiation): three Fp4 squarings = 9 Fp2 squarings.
HD void fp4_sqr(const fp2& a, const fp2& b, fp2& c0, fp2& c1) {
    fp2 t0 = fp2_sqr(a);
    fp2 t1 = fp2_sqr(b);
    c0 = fp2_add(fp2_mul_xi(t1), t0);
    c1 = fp2_sub(fp2_sub(fp2_sqr(fp2_add(a, b)), t0), t1);
}
HDN fp12 fp12_cyclotomic_sqr(const fp12& x) {
    fp2 z0 = x.b0.a0, z4 = x.b0.a1, z3 = x.b0.a2, z2 = x.b1.a0, z1 = x.b1.a1, z5 = x.b1.a2;
    fp2 t0, t1, t2, t3;
    fp4_sqr(z0, z1, t0, t1);
    z0 = fp2_add(fp2_dbl(fp2_sub(t0, z0)), t0);
    z1 = fp2_add(fp2_dbl(fp2_add(t1, z1)), t1);
    fp4_sqr(z2, z3, t0, t1);
    fp4_sqr(z4, z5, t2, t3);
    z4 = fp2_add(fp2_dbl(fp2_sub(t0, z4)), t0);
    z5 = fp2_add(fp2_dbl(fp2_add(t1, z5)), t1);
    fp2 t = fp2_mul_xi(t3);
    z2 = fp2_add(fp2_dbl(fp2_add(t, z2)), t);
    z3 = fp2_add(fp2_dbl(fp2_sub(t2, z3)), t2);
    fp12 r;
    r.b0.a0 = z0;
    r.b0.a1 = z4;
    r.b0.a2 = z3;
    r.b1.a0 = z2;
    r.b1.a1 = z1;
    r.b1.a2 = z5;
    return r;
}

}  // namespace b2
