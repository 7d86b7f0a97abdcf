// team.cuh -- lane-cooperative ("team") versions of the pairing for the latency-bound verification tail.
//
// With one thread per aggregate, 2 048 aggregates are only 64 warps: each warp's serial chain of IMAD.WIDE
// (one per 4 cycles on its scheduler's fmaheavy pipe) sets the kernel time while 90 % of the chip idles.  SIMT lanes
// are free, so three lanes cooperate on one Fp12 value: the three Fp6 products of a Karatsuba Fp12 multiplication,
// the three Fp4 squarings of a Granger-Scott cyclotomic squaring, or three independent Fp2 products of a G2 point
// doubling run in the SAME warp instruction stream on different lanes.  Operands live in shared memory (a `team_ws`
// per team); every phase is "all lanes read inputs, compute one sub-product with ONE call site (no divergence),
// write disjoint outputs", separated by team_sync().  Critical path per Fp12 multiplication: 18 Fp mul instead of 54;
// cyclotomic squaring 6 instead of 18; Miller doubling step ~12 instead of ~33.
//
// The code is host+device: tests/hostsim runs a team as three host threads with a barrier for team_sync(), and
// checks the result against the oracle (and against the single-thread pairing.cuh) bit for bit.
#pragma once
#include "pairing.cuh"

namespace b2 {

struct team {
    int lane;            // 0..2
    unsigned mask;       // device: lanes of the warp that take part in team_sync (all teams of the warp sync together)
    void* host_barrier;  // host-sim only
};

#if defined(__CUDA_ARCH__)
DEV void team_sync(const team& t) { __syncwarp(t.mask); }
#else
void host_team_barrier(void* b);    // provided by tests/hostsim (pthread barrier)
inline void team_sync(const team& t) { host_team_barrier(t.host_barrier); }
#endif

struct team_ws {
    fp12 f, g, h, m;     // accumulators / temporaries of the final exponentiation
    fp6 t[3];            // Karatsuba partial products
    fp2 u[12];           // Fp2 temporaries of the Miller steps
    g2_jac T;            // running point of the Miller loop
    line_coeffs line;
};

HD fp2* fp6_coeff(fp6* x, int k) { return (&x->a0) + k; }
HD const fp2* fp6_coeff(const fp6* x, int k) { return (&x->a0) + k; }
// coefficient of w^k, k = 0..5  (w^0..w^5 = b0.a0, b1.a0, b0.a1, b1.a1, b0.a2, b1.a2)
HD fp2* fp12_wcoeff(fp12* x, int k) { return fp6_coeff((k & 1) ? &x->b1 : &x->b0, k >> 1); }
HD const fp2* fp12_wcoeff(const fp12* x, int k) { return fp6_coeff((k & 1) ? &x->b1 : &x->b0, k >> 1); }

// out.b0 = t0 + v*t1 ; out.b1 = t2 - t0 - t1 ; lane k writes coefficient k of both halves
HD void team_karatsuba_combine(const team& tm, fp12* out, const fp6* t) {
    const int k = tm.lane;
    fp2 t0k = *fp6_coeff(&t[0], k), t1k = *fp6_coeff(&t[1], k), t2k = *fp6_coeff(&t[2], k);
    fp2 v1 = (k == 0) ? fp2_mul_xi(t[1].a2) : *fp6_coeff(&t[1], k - 1 < 0 ? 0 : k - 1);
    *fp6_coeff(&out->b0, k) = fp2_add(t0k, v1);
    *fp6_coeff(&out->b1, k) = fp2_sub(fp2_sub(t2k, t0k), t1k);
}

// out = a * b   (out may alias a or b)
HD void team_fp12_mul(const team& tm, team_ws* ws, fp12* out, const fp12* a, const fp12* b) {
    fp6 x = a->b0, y = b->b0;
    if (tm.lane == 1) {
        x = a->b1;
        y = b->b1;
    } else if (tm.lane == 2) {
        x = fp6_add(a->b0, a->b1);
        y = fp6_add(b->b0, b->b1);
    }
    fp6 r = fp6_mul(x, y);
    ws->t[tm.lane] = r;
    team_sync(tm);
    team_karatsuba_combine(tm, out, ws->t);
    team_sync(tm);
}

// out = a^2 (general element), complex squaring: t = a0*a1 (lane 0), s = (a0+a1)(a0+v*a1) (lane 1); lane 2 mirrors lane 0
HD void team_fp12_sqr(const team& tm, team_ws* ws, fp12* out, const fp12* a) {
    fp6 x = a->b0, y = a->b1;
    if (tm.lane == 1) {
        x = fp6_add(a->b0, a->b1);
        y = fp6_add(a->b0, fp6_mul_v(a->b1));
    }
    fp6 r = fp6_mul(x, y);
    ws->t[tm.lane] = r;
    team_sync(tm);
    const int k = tm.lane;
    fp2 tk = *fp6_coeff(&ws->t[0], k), sk = *fp6_coeff(&ws->t[1], k);
    fp2 vt = (k == 0) ? fp2_mul_xi(ws->t[0].a2) : *fp6_coeff(&ws->t[0], k - 1 < 0 ? 0 : k - 1);
    fp2 o0 = fp2_sub(fp2_sub(sk, tk), vt), o1 = fp2_dbl(tk);
    team_sync(tm);                                   // every lane has read ws->t before `out` (possibly == a) changes
    *fp6_coeff(&out->b0, k) = o0;
    *fp6_coeff(&out->b1, k) = o1;
    team_sync(tm);
}

// out = a * ((l0 + l1 v) + (l4 v) w)
HD void team_fp12_mul_by_014(const team& tm, team_ws* ws, fp12* out, const fp12* a, const fp2& l0, const fp2& l1, const fp2& l4) {
    fp6 x = a->b0;
    fp2 c0 = l0, c1 = l1;
    if (tm.lane == 1) {
        x = a->b1;
        c0 = fp2_zero();
        c1 = l4;
    } else if (tm.lane == 2) {
        x = fp6_add(a->b0, a->b1);
        c1 = fp2_add(l1, l4);
    }
    fp6 r = fp6_mul_by_01(x, c0, c1);
    ws->t[tm.lane] = r;
    team_sync(tm);
    team_karatsuba_combine(tm, out, ws->t);
    team_sync(tm);
}

// Granger-Scott cyclotomic squaring: lane k squares the Fp4 element (g_k, g_{k+3}); outputs as in fp12_cyclotomic_sqr
HD void team_fp12_cyc_sqr(const team& tm, fp12* out, const fp12* a) {
    const int k = tm.lane;
    fp2 za = *fp12_wcoeff(a, k), zb = *fp12_wcoeff(a, k + 3);
    fp2 t0, t1;
    fp4_sqr(za, zb, t0, t1);
    // lane 0 updates (g0, g3) = (z0, z1); lane 1 updates (g2, g5) = (z4, z5); lane 2 updates (g1, g4) = (z2, z3)
    const int ia = (k == 0) ? 0 : (k == 1 ? 2 : 1);
    fp2 ra = *fp12_wcoeff(a, ia), rb = *fp12_wcoeff(a, ia + 3);
    fp2 oa, ob;
    if (k < 2) {
        oa = fp2_add(fp2_dbl(fp2_sub(t0, ra)), t0);
        ob = fp2_add(fp2_dbl(fp2_add(t1, rb)), t1);
    } else {
        fp2 u = fp2_mul_xi(t1);
        oa = fp2_add(fp2_dbl(fp2_add(u, ra)), u);
        ob = fp2_add(fp2_dbl(fp2_sub(t0, rb)), t0);
    }
    team_sync(tm);
    *fp12_wcoeff(out, ia) = oa;
    *fp12_wcoeff(out, ia + 3) = ob;
    team_sync(tm);
}

HD void team_fp12_conj(const team& tm, fp12* out, const fp12* a) {
    const int k = tm.lane;
    fp2 c0 = *fp6_coeff(&a->b0, k), c1 = fp2_neg(*fp6_coeff(&a->b1, k));
    team_sync(tm);
    *fp6_coeff(&out->b0, k) = c0;
    *fp6_coeff(&out->b1, k) = c1;
    team_sync(tm);
}
// p-power Frobenius: g_k -> conj(g_k) * gamma^k ; lane k handles k and k+3
HD void team_fp12_frob(const team& tm, fp12* out, const fp12* a) {
    const int k = tm.lane;
    fp2 lo = fp2_conj(*fp12_wcoeff(a, k)), hi = fp2_conj(*fp12_wcoeff(a, k + 3));
    fp2 clo = (k == 0) ? fp2_one() : fp2_load_const(C_FROB1_1 + 24 * (k - 1));
    fp2 chi = fp2_load_const(C_FROB1_1 + 24 * (k + 2));
    lo = fp2_mul(lo, clo);
    hi = fp2_mul(hi, chi);
    team_sync(tm);
    *fp12_wcoeff(out, k) = lo;
    *fp12_wcoeff(out, k + 3) = hi;
    team_sync(tm);
}
// p^2-power Frobenius: g_k -> g_k * norm(gamma^k) (in Fp)
HD void team_fp12_frob2(const team& tm, fp12* out, const fp12* a) {
    const int k = tm.lane;
    fp clo = (k == 0) ? fp_one() : fp_load_const(C_FROB2_1 + 12 * (k - 1));
    fp chi = fp_load_const(C_FROB2_1 + 12 * (k + 2));
    fp2 lo = fp2_mul_fp(*fp12_wcoeff(a, k), clo), hi = fp2_mul_fp(*fp12_wcoeff(a, k + 3), chi);
    team_sync(tm);
    *fp12_wcoeff(out, k) = lo;
    *fp12_wcoeff(out, k + 3) = hi;
    team_sync(tm);
}

// r = a^|x| for a in the cyclotomic subgroup; r must not alias a
HD void team_cyc_exp_x_abs(const team& tm, team_ws* ws, fp12* r, const fp12* a) {
    if (tm.lane == 0) *r = *a;
    team_sync(tm);
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        team_fp12_cyc_sqr(tm, r, r);
        if ((B2_X_ABS >> i) & 1ull) team_fp12_mul(tm, ws, r, r, a);
    }
}
// r = a^x (x < 0)
HD void team_cyc_exp_x(const team& tm, team_ws* ws, fp12* r, const fp12* a) {
    team_cyc_exp_x_abs(tm, ws, r, a);
    team_fp12_conj(tm, r, r);
}

// ws->f := final_exponentiation(ws->f)   (same chain as pairing.cuh)
HD void team_final_exponentiation(const team& tm, team_ws* ws) {
    fp12 *f = &ws->f, *g = &ws->g, *h = &ws->h, *m = &ws->m;
    if (tm.lane == 0) *g = fp12_inv(*f);              // one Fp inversion: serial
    team_sync(tm);
    team_fp12_conj(tm, h, f);
    team_fp12_mul(tm, ws, h, h, g);                   // t = conj(f) / f
    team_fp12_frob2(tm, g, h);
    team_fp12_mul(tm, ws, m, g, h);                   // m = t^(p^2+1)
    // a = m^((x-1)^2)
    team_cyc_exp_x(tm, ws, g, m);
    team_fp12_conj(tm, h, m);
    team_fp12_mul(tm, ws, f, g, h);                   // f = m^(x-1)
    team_cyc_exp_x(tm, ws, g, f);
    team_fp12_conj(tm, h, f);
    team_fp12_mul(tm, ws, f, g, h);                   // f = a
    // b = a^(x+p)
    team_cyc_exp_x(tm, ws, g, f);
    team_fp12_frob(tm, h, f);
    team_fp12_mul(tm, ws, f, g, h);                   // f = b
    // c = b^(x^2 + p^2 - 1)
    team_cyc_exp_x(tm, ws, g, f);
    team_cyc_exp_x(tm, ws, h, g);                     // h = b^(x^2)
    team_fp12_frob2(tm, g, f);
    team_fp12_mul(tm, ws, h, h, g);
    team_fp12_conj(tm, g, f);
    team_fp12_mul(tm, ws, f, h, g);                   // f = c
    // result = c * m^3
    team_fp12_cyc_sqr(tm, g, m);
    team_fp12_mul(tm, ws, g, g, m);
    team_fp12_mul(tm, ws, f, f, g);
}

// ------------------------------------------------------------------------------------------ Miller loop, 3 lanes
// Doubling step in five rounds of (at most) three independent Fp2 products; same formulas as miller_dbl_step.
HD void team_miller_dbl_step(const team& tm, team_ws* ws, const miller_p& P) {
    g2_jac* T = &ws->T;
    fp2* u = ws->u;
    const int k = tm.lane;
    // R1: A = X^2, B = Y^2, ZZ = Z^2
    u[k] = fp2_sqr(k == 0 ? T->x : (k == 1 ? T->y : T->z));
    team_sync(tm);
    // R2: C = B^2, XB = (X+B)^2, F = (3A)^2
    {
        fp2 in = (k == 0) ? u[1] : (k == 1 ? fp2_add(T->x, u[1]) : fp2_mul3(u[0]));
        u[3 + k] = fp2_sqr(in);
    }
    team_sync(tm);
    fp2 A = u[0], B = u[1], ZZ = u[2], C = u[3], XB = u[4], F = u[5];
    fp2 E = fp2_mul3(A);
    fp2 D = fp2_dbl(fp2_sub(fp2_sub(XB, A), C));
    fp2 X3 = fp2_sub(F, fp2_dbl(D));
    // R3: M1 = E*(D - X3), YZ = Y*Z, EX = E*X
    {
        fp2 a = (k == 1) ? T->y : E;
        fp2 b = (k == 0) ? fp2_sub(D, X3) : (k == 1 ? T->z : T->x);
        u[6 + k] = fp2_mul(a, b);
    }
    team_sync(tm);
    fp2 Z3 = fp2_dbl(u[7]);
    // R4: EZZ = E*ZZ, ZZZ = Z3*ZZ   (lane 2 mirrors lane 0)
    {
        fp2 a = (k == 1) ? Z3 : E;
        u[9 + k] = fp2_mul(a, ZZ);
    }
    team_sync(tm);
    // R5: scale by the three Fp factors of P and store
    {
        fp2 v = (k == 0) ? fp2_sub(u[8], fp2_dbl(B)) : (k == 1 ? u[9] : u[10]);
        fp s = (k == 0) ? P.k0 : (k == 1 ? P.k1 : P.k3);
        fp2 o = fp2_mul_fp(v, s);
        fp2 y3 = fp2_sub(u[6], fp2_mul8(C));
        team_sync(tm);
        if (k == 0) {
            ws->line.c0 = o;
            T->x = X3;
        } else if (k == 1) {
            ws->line.c1 = o;
            T->y = y3;
        } else {
            ws->line.d1 = o;
            T->z = Z3;
        }
    }
    team_sync(tm);
}

// f_{|x|,Q}(P) conjugated, into ws->f.  The five addition steps run on lane 0 (negligible).
HD void team_miller_loop(const team& tm, team_ws* ws, const g1_jac& Pj, const g2_aff& Q, bool q_inf) {
    if (q_inf || pt_is_inf(Pj)) {
        if (tm.lane == 0) ws->f = fp12_one();
        team_sync(tm);
        return;
    }
    miller_p P = miller_p_from_jac(Pj);
    if (tm.lane == 0) {
        ws->T = pt_from_affine(Q);
        ws->f = fp12_one();
    }
    team_sync(tm);
#pragma unroll 1
    for (int i = 62; i >= 0; i--) {
        team_fp12_sqr(tm, ws, &ws->f, &ws->f);
        team_miller_dbl_step(tm, ws, P);
        team_fp12_mul_by_014(tm, ws, &ws->f, &ws->f, ws->line.c0, ws->line.c1, ws->line.d1);
        if ((B2_X_ABS >> i) & 1ull) {
            if (tm.lane == 0) {
                g2_jac T = ws->T;
                line_coeffs l;
                miller_add_step(T, Q, P, l);
                ws->T = T;
                ws->line = l;
            }
            team_sync(tm);
            team_fp12_mul_by_014(tm, ws, &ws->f, &ws->f, ws->line.c0, ws->line.c1, ws->line.d1);
        }
    }
    team_fp12_conj(tm, &ws->f, &ws->f);
}

}  // namespace b2
