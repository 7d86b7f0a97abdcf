// h2c.cuh -- SHA-256, expand_message_xmd and hash_to_curve for BLS12381G2_XMD:SHA-256_SSWU_RO_
// (RFC 9380 sections 5.2, 5.3.1, 6.6.2, 8.8.2, appendix E.3 and G.3).  Kernel K4 of SURVEY.md section 2.
// The reference reaches this only through is_valid_indexed_attestation
// (/root/reference/pos-evolution.md:736, :976) -> bls.FastAggregateVerify -> hash_to_G2.
//
// Per message: 2 x map_to_curve, each with exactly three Fp exponentiations and no
// data-dependent branch over them (inverse, norm root, Fp2 root -- see sswu_map), one
// inversion-free 3-isogeny to Jacobian coordinates, one addition, and psi-based cofactor clearing.
#pragma once
#include "curve.cuh"

namespace b2 {

// ------------------------------------------------------------------------------------------ SHA-256
struct sha256_ctx {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t fill;
    uint64_t total;
};
HD uint32_t sha_rotr(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }
HD uint32_t sha_k(int i) {
    constexpr uint32_t K[64] = {
        0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01, 0x243185be,
        0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa,
        0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967, 0x27b70a85,
        0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85, 0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3,
        0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f,
        0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
    return K[i];
}
HDN void sha256_compress(uint32_t* h, const uint8_t* blk) {
    uint32_t w[64];
#pragma unroll
    for (int i = 0; i < 16; i++)
        w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
#pragma unroll
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = sha_rotr(w[i - 15], 7) ^ sha_rotr(w[i - 15], 18) ^ (w[i - 15] >> 3);
        uint32_t s1 = sha_rotr(w[i - 2], 17) ^ sha_rotr(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        uint32_t S1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + sha_k(i) + w[i];
        uint32_t S0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
HD void sha256_init(sha256_ctx& c) {
    c.h[0] = 0x6a09e667; c.h[1] = 0xbb67ae85; c.h[2] = 0x3c6ef372; c.h[3] = 0xa54ff53a;
    c.h[4] = 0x510e527f; c.h[5] = 0x9b05688c; c.h[6] = 0x1f83d9ab; c.h[7] = 0x5be0cd19;
    c.fill = 0;
    c.total = 0;
}
HD void sha256_update(sha256_ctx& c, const uint8_t* p, uint32_t n) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        c.buf[c.fill++] = p[i];
        if (c.fill == 64) {
            sha256_compress(c.h, c.buf);
            c.fill = 0;
        }
    }
    c.total += n;
}
HD void sha256_final(sha256_ctx& c, uint8_t* out) {
    uint64_t bits = c.total * 8;
    uint8_t pad = 0x80;
    sha256_update(c, &pad, 1);
    pad = 0;
    while (c.fill != 56) sha256_update(c, &pad, 1);
    uint8_t len[8];
#pragma unroll
    for (int i = 0; i < 8; i++) len[i] = (uint8_t)(bits >> (56 - 8 * i));
    sha256_update(c, len, 8);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        out[4 * i] = (uint8_t)(c.h[i] >> 24);
        out[4 * i + 1] = (uint8_t)(c.h[i] >> 16);
        out[4 * i + 2] = (uint8_t)(c.h[i] >> 8);
        out[4 * i + 3] = (uint8_t)c.h[i];
    }
}

// expand_message_xmd(msg, DST, 256) -> 8 blocks of 32 bytes (RFC 9380 section 5.3.1)
HDN void expand_message_xmd_256(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len, uint8_t* out) {
    uint8_t b0[32], bi[32], tmp[32];
    uint8_t dlen = (uint8_t)dst_len;
    sha256_ctx c;
    sha256_init(c);
    uint8_t z = 0;
#pragma unroll 1
    for (int i = 0; i < 64; i++) sha256_update(c, &z, 1);
    sha256_update(c, msg, msg_len);
    uint8_t lib[3] = {1, 0, 0};                 // len_in_bytes = 256 as two bytes, then the 0x00 counter
    sha256_update(c, lib, 3);
    sha256_update(c, dst, dst_len);
    sha256_update(c, &dlen, 1);
    sha256_final(c, b0);
#pragma unroll 1
    for (int blk = 1; blk <= 8; blk++) {
        sha256_init(c);
        if (blk == 1) {
            sha256_update(c, b0, 32);
        } else {
            for (int k = 0; k < 32; k++) tmp[k] = b0[k] ^ bi[k];
            sha256_update(c, tmp, 32);
        }
        uint8_t ctr = (uint8_t)blk;
        sha256_update(c, &ctr, 1);
        sha256_update(c, dst, dst_len);
        sha256_update(c, &dlen, 1);
        sha256_final(c, bi);
        for (int k = 0; k < 32; k++) out[(blk - 1) * 32 + k] = bi[k];
    }
}

// 64 big-endian bytes -> Fp (Montgomery): (hi * 2^256 + lo) mod p
HD fp fp_from_be64_reduce(const uint8_t* in) {
    fp hi = fp_zero(), lo = fp_zero();
#pragma unroll
    for (int i = 0; i < 8; i++) {
        hi.l[7 - i] = ((uint32_t)in[4 * i] << 24) | ((uint32_t)in[4 * i + 1] << 16) | ((uint32_t)in[4 * i + 2] << 8) | in[4 * i + 3];
        lo.l[7 - i] = ((uint32_t)in[32 + 4 * i] << 24) | ((uint32_t)in[32 + 4 * i + 1] << 16) | ((uint32_t)in[32 + 4 * i + 2] << 8) | in[32 + 4 * i + 3];
    }
    return fp_add(fp_mul(fp_to_mont(hi), fp_load_const(C_TWO256)), fp_to_mont(lo));
}

// hash_to_field with m = 2, count = 2 (RFC 9380 section 5.2)
HDN void hash_to_field_fp2x2(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len, fp2& u0, fp2& u1) {
    uint8_t uniform[256];
    expand_message_xmd_256(msg, msg_len, dst, dst_len, uniform);
    u0.c0 = fp_from_be64_reduce(uniform);
    u0.c1 = fp_from_be64_reduce(uniform + 64);
    u1.c0 = fp_from_be64_reduce(uniform + 128);
    u1.c1 = fp_from_be64_reduce(uniform + 192);
}

// ------------------------------------------------------------------------------------------ SSWU
// Second half of the Fp2 square root: given s with s^2 = norm(a) and a.c1 != 0, one exponentiation.
HDN fp2 fp2_sqrt_given_norm_root(const fp2& a, const fp& s) {
    fp half = fp_load_const(C_TWO_INV);
    fp t = fp_mul(fp_add(a.c0, s), half);
    fp d = fp_pow_pm3d4(t);
    fp x = fp_mul(d, t);
    fp y = fp_mul(fp_mul(a.c1, half), d);
    bool direct = fp_eq(fp_sqr(x), t);
    fp2 r;
    r.c0 = fp_select(direct, x, y);
    r.c1 = fp_select(direct, y, fp_neg(x));
    return r;
}

// Simplified SWU for AB != 0 onto E2': y^2 = x^3 + A'x + B' (RFC 9380 section 6.6.2).  Exactly three Fp
// exponentiations for every input:
//   (1) 1/tv1                                   -> x1, gx1, and x2 = Z u^2 x1, gx2 = (Z u^2)^3 gx1
//   (2) s = norm(gx1)^((p+1)/4).  gx1 is a square in Fp2 iff s^2 == norm(gx1).  Otherwise
//       s^2 = -norm(gx1) and sqrt(norm(gx2)) = norm(Z) * norm(u)^3 * zeta * s with
//       zeta = sqrt(-norm(Z)) -- no second norm exponentiation is needed;
//   (3) the Fp2 root of whichever of gx1 / gx2 is the square.
HDN void sswu_map(const fp2& u, fp2& x_out, fp2& y_out) {
    const fp2 A = fp2_load_const(C_SSWU_A), B = fp2_load_const(C_SSWU_B), Z = fp2_load_const(C_SSWU_Z);
    fp2 u2 = fp2_sqr(u);
    fp2 zu2 = fp2_mul(Z, u2);
    fp2 tv1 = fp2_add(fp2_sqr(zu2), zu2);
    fp2 x1;
    if (fp2_is_zero(tv1)) {
        x1 = fp2_load_const(C_SSWU_B_OVER_ZA);
    } else {
        x1 = fp2_mul(fp2_load_const(C_SSWU_MB_OVER_A), fp2_add(fp2_one(), fp2_inv(tv1)));
    }
    fp2 gx1 = fp2_add(fp2_add(fp2_mul(fp2_sqr(x1), x1), fp2_mul(A, x1)), B);
    fp2 x2 = fp2_mul(zu2, x1);
    fp2 gx2 = fp2_mul(fp2_mul(fp2_sqr(zu2), zu2), gx1);
    fp n1 = fp_add(fp_sqr(gx1.c0), fp_sqr(gx1.c1));
    fp s1 = fp_mul(fp_pow_pm3d4(n1), n1);
    bool is_sq = fp_eq(fp_sqr(s1), n1);
    fp nu = fp_add(fp_sqr(u.c0), fp_sqr(u.c1));
    fp s2 = fp_mul(fp_mul(fp_mul(fp_sqr(nu), nu), fp_mul(fp_load_const(C_SSWU_ZNORM), fp_load_const(C_SSWU_ZETA))), s1);
    fp2 g = fp2_select(is_sq, gx1, gx2);
    fp s = fp_select(is_sq, s1, s2);
    fp2 x = fp2_select(is_sq, x1, x2);
    fp2 y;
    if (fp_is_zero(g.c1)) {
        fp2_sqrt(g, y);                          // measure-zero input class; generic path
    } else {
        y = fp2_sqrt_given_norm_root(g, s);
    }
    if (fp2_sgn0(u) != fp2_sgn0(y)) y = fp2_neg(y);
    x_out = x;
    y_out = y;
}

// 3-isogeny E2' -> E2 (RFC 9380 appendix E.3), affine in, Jacobian out, no inversion:
//   x = xn/xd, y = y' * yn/yd;  with Z = xd*yd:  X = xn*xd*yd^2,  Y = y'*yn*xd^3*yd^2.
HD fp2 iso_horner(const fp2& x, int off, int ncoef, bool monic) {
    fp2 acc = monic ? fp2_add(x, fp2_load_const(off + 24 * (ncoef - 1))) : fp2_load_const(off + 24 * (ncoef - 1));
#pragma unroll 1
    for (int i = ncoef - 2; i >= 0; i--) acc = fp2_add(fp2_mul(acc, x), fp2_load_const(off + 24 * i));
    return acc;
}
HDN g2_jac iso3_map(const fp2& x, const fp2& y) {
    fp2 xn = iso_horner(x, C_ISO_XNUM0, 4, false);
    fp2 xd = iso_horner(x, C_ISO_XDEN0, 2, true);
    fp2 yn = iso_horner(x, C_ISO_YNUM0, 4, false);
    fp2 yd = iso_horner(x, C_ISO_YDEN0, 3, true);
    g2_jac r;
    r.z = fp2_mul(xd, yd);                       // zero <=> the image is the point at infinity
    fp2 yd2 = fp2_sqr(yd);
    fp2 t = fp2_mul(xd, yd2);                    // xd * yd^2
    r.x = fp2_mul(xn, t);
    r.y = fp2_mul(fp2_mul(y, yn), fp2_mul(fp2_sqr(xd), t));
    return r;
}

// clear_cofactor_bls12381_g2 (RFC 9380 appendix G.3) == multiplication by h_eff
HDN g2_jac g2_clear_cofactor(const g2_jac& p) {
    g2_jac t1 = pt_neg(pt_mul_u64(p, B2_X_ABS));          // c1 * P, c1 = x < 0
    g2_jac t2 = g2_psi(p);
    g2_jac t3 = g2_psi2(pt_dbl(p));
    t3 = pt_add(t3, pt_neg(t2));
    t2 = pt_add(t1, t2);
    t2 = pt_neg(pt_mul_u64(t2, B2_X_ABS));
    t3 = pt_add(t3, t2);
    t3 = pt_add(t3, pt_neg(t1));
    return pt_add(t3, pt_neg(p));
}

HD g2_jac map_to_curve_g2(const fp2& u) {
    fp2 x, y;
    sswu_map(u, x, y);
    return iso3_map(x, y);
}

// hash_to_curve -> Jacobian point in G2
HDN g2_jac hash_to_g2(const uint8_t* msg, uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
    fp2 u0, u1;
    hash_to_field_fp2x2(msg, msg_len, dst, dst_len, u0, u1);
    g2_jac q = pt_add(map_to_curve_g2(u0), map_to_curve_g2(u1));
    return g2_clear_cofactor(q);
}

}  // namespace b2
