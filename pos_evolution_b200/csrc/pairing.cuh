// pairing.cuh -- optimal ate pairing on BLS12-381: Miller loop (kernel K5) and final
// exponentiation (kernel K6) of SURVEY.md section 2.  Reference call site: bls.FastAggregateVerify inside
// is_valid_indexed_attestation (/root/reference/pos-evolution.md:736, :976); the arithmetic is
// third-party there (py_ecc) and restated in oracle/bls12_381.py, whose tower and final-exponent
// (3*(p^12-1)/r) this file matches so that GT values compare bit for bit.
//
// Untwist (x', y') -> (x'/w^2, y'/w^3).  The line through T (tangent or chord) evaluated at
// P = (xP, yP), times w^3 and any Fp2 factor (all killed by the final exponentiation), is
//        c0 + c1 * w^2 + d1 * w^3      <->  tower slots  b0.a0, b0.a1, b1.a1   ("014").
// T is kept in Jacobian coordinates; P may be Jacobian too (XP, YP, ZP): the line is then
// additionally scaled by ZP^3, which avoids the field inversion for the aggregated pubkey.
#pragma once
#include "curve.cuh"
#include "fp12.cuh"

namespace b2 {

struct line_coeffs {
    fp2 c0, c1, d1;
};

// the three Fp factors by which P enters the lines
struct miller_p {
    fp k0;   // multiplies c0 :  ZP^3        (1 for affine P)
    fp k1;   // multiplies c1 :  -XP * ZP    (-xP)
    fp k3;   // multiplies d1 :  YP          (yP)
};
HD miller_p miller_p_from_jac(const g1_jac& p) {
    miller_p m;
    fp z2 = fp_sqr(p.z);
    m.k0 = fp_mul(z2, p.z);
    m.k1 = fp_neg(fp_mul(p.x, p.z));
    m.k3 = p.y;
    return m;
}

// tangent at T (Jacobian), then T <- 2T.   slope = 3X^2 / (2YZ); scaled by 2YZ^3 = Z3 * Z^2:
//   c0 = 3X^3 - 2Y^2,  c1 = -3X^2 Z^2 * xP,  d1 = Z3 Z^2 * yP
HDN void miller_dbl_step(g2_jac& T, const miller_p& P, line_coeffs& l) {
    fp2 A = fp2_sqr(T.x);
    fp2 B = fp2_sqr(T.y);
    fp2 C = fp2_sqr(B);
    fp2 D = fp2_dbl(fp2_sub(fp2_sub(fp2_sqr(fp2_add(T.x, B)), A), C));
    fp2 E = fp2_mul3(A);
    fp2 ZZ = fp2_sqr(T.z);
    g2_jac r;
    r.x = fp2_sub(fp2_sqr(E), fp2_dbl(D));
    r.y = fp2_sub(fp2_mul(E, fp2_sub(D, r.x)), fp2_mul8(C));
    r.z = fp2_dbl(fp2_mul(T.y, T.z));
    l.c0 = fp2_mul_fp(fp2_sub(fp2_mul(E, T.x), fp2_dbl(B)), P.k0);
    l.c1 = fp2_mul_fp(fp2_mul(E, ZZ), P.k1);
    l.d1 = fp2_mul_fp(fp2_mul(r.z, ZZ), P.k3);
    T = r;
}

// chord through T (Jacobian) and Q (affine), then T <- T + Q.  slope = (S2 - Y)/(Z H); scaled by Z3 = 2ZH:
//   c0 = r*xQ - yQ*Z3,  c1 = -r * xP,  d1 = Z3 * yP       (r = 2(S2 - Y))
HDN void miller_add_step(g2_jac& T, const g2_aff& Q, const miller_p& P, line_coeffs& l) {
    fp2 Z1Z1 = fp2_sqr(T.z);
    fp2 U2 = fp2_mul(Q.x, Z1Z1);
    fp2 S2 = fp2_mul(fp2_mul(Q.y, T.z), Z1Z1);
    fp2 H = fp2_sub(U2, T.x);
    fp2 HH = fp2_sqr(H);
    fp2 I = fp2_mul4(HH);
    fp2 J = fp2_mul(H, I);
    fp2 rr = fp2_dbl(fp2_sub(S2, T.y));
    fp2 V = fp2_mul(T.x, I);
    g2_jac r;
    r.x = fp2_sub(fp2_sub(fp2_sqr(rr), J), fp2_dbl(V));
    r.y = fp2_sub(fp2_mul(rr, fp2_sub(V, r.x)), fp2_dbl(fp2_mul(T.y, J)));
    r.z = fp2_sub(fp2_sub(fp2_sqr(fp2_add(T.z, H)), Z1Z1), HH);
    l.c0 = fp2_mul_fp(fp2_sub(fp2_mul(rr, Q.x), fp2_mul(Q.y, r.z)), P.k0);
    l.c1 = fp2_mul_fp(rr, P.k1);
    l.d1 = fp2_mul_fp(r.z, P.k3);
    T = r;
}

// f_{|x|,Q}(P), conjugated because x < 0.  Either argument at infinity -> 1 (as the oracle).
HDN fp12 miller_loop(const g1_jac& Pj, const g2_aff& Q, bool q_inf) {
    if (q_inf || pt_is_inf(Pj)) return fp12_one();
    miller_p P = miller_p_from_jac(Pj);
    g2_jac T = pt_from_affine(Q);
    fp12 f = fp12_one();
    line_coeffs l;
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        f = fp12_sqr(f);
        miller_dbl_step(T, P, l);
        f = fp12_mul_by_014(f, l.c0, l.c1, l.d1);
        if ((B2_X_ABS >> i) & 1ull) {
            miller_add_step(T, Q, P, l);
            f = fp12_mul_by_014(f, l.c0, l.c1, l.d1);
        }
    }
    return fp12_conj(f);
}

// a^|x| for a in the cyclotomic subgroup
HDN fp12 fp12_cyc_exp_x_abs(const fp12& a) {
    fp12 r = a;
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        r = fp12_cyclotomic_sqr(r);
        if ((B2_X_ABS >> i) & 1ull) r = fp12_mul(r, a);
    }
    return r;
}
// a^x (x negative): inverse == conjugate in the cyclotomic subgroup
HD fp12 fp12_cyc_exp_x(const fp12& a) { return fp12_conj(fp12_cyc_exp_x_abs(a)); }

// f^(3*(p^12-1)/r): easy part (p^6-1)(p^2+1), hard part by the Hayashida-Hayasaka-Teruya chain
//   3*Phi12(p)/r = (x-1)^2 (x+p) (x^2+p^2-1) + 3
HDN fp12 final_exponentiation(const fp12& f) {
    fp12 t = fp12_mul(fp12_conj(f), fp12_inv(f));
    fp12 m = fp12_mul(fp12_frob2(t), t);
    fp12 a = fp12_mul(fp12_cyc_exp_x(m), fp12_conj(m));
    a = fp12_mul(fp12_cyc_exp_x(a), fp12_conj(a));
    fp12 b = fp12_mul(fp12_cyc_exp_x(a), fp12_frob(a));
    fp12 c = fp12_mul(fp12_mul(fp12_cyc_exp_x(fp12_cyc_exp_x(b)), fp12_frob2(b)), fp12_conj(b));
    return fp12_mul(c, fp12_mul(fp12_cyclotomic_sqr(m), m));
}

}  // namespace b2
