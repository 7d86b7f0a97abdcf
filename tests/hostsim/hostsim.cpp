// hostsim.cpp -- TEST-ONLY host build of the device math headers (pos_evolution_b200/csrc/*.cuh).
// Compiled with g++ into tests/hostsim/libhostsim.so by tests/hostsim/build.py; the carry-chain
// primitives are emulated (csrc/platform.cuh), everything above them is the very code the CUDA
// kernels run.  Lets the CPU test-suite check field towers, curve ops, hash-to-curve and the
// pairing against the oracle without a GPU.  Never linked into, or loaded by, the product.
#include <string.h>

#include "cores.cuh"

using namespace b2;

template <class T> static T ld(const uint32_t* p) {
    T v;
    memcpy(&v, p, sizeof(T));
    return v;
}
template <class T> static void st(uint32_t* p, const T& v) { memcpy(p, &v, sizeof(T)); }

extern "C" {

// ---- Fp / Fp2 (Montgomery limbs in, Montgomery limbs out)
void hs_fp_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fp_mul(ld<fp>(a), ld<fp>(b))); }
void hs_fp_add(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fp_add(ld<fp>(a), ld<fp>(b))); }
void hs_fp_sub(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fp_sub(ld<fp>(a), ld<fp>(b))); }
void hs_fp_neg(const uint32_t* a, uint32_t* r) { st(r, fp_neg(ld<fp>(a))); }
void hs_fp_inv(const uint32_t* a, uint32_t* r) { st(r, fp_inv(ld<fp>(a))); }
int hs_fp_sqrt(const uint32_t* a, uint32_t* r) {
    fp o = fp_zero();
    bool ok = fp_sqrt(ld<fp>(a), o);
    st(r, o);
    return ok;
}
void hs_fp_to_mont(const uint32_t* a, uint32_t* r) { st(r, fp_to_mont(ld<fp>(a))); }
void hs_fp_from_mont(const uint32_t* a, uint32_t* r) { st(r, fp_from_mont(ld<fp>(a))); }
void hs_fp2_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fp2_mul(ld<fp2>(a), ld<fp2>(b))); }
void hs_fp2_sqr(const uint32_t* a, uint32_t* r) { st(r, fp2_sqr(ld<fp2>(a))); }
void hs_fp2_inv(const uint32_t* a, uint32_t* r) { st(r, fp2_inv(ld<fp2>(a))); }
int hs_fp2_sqrt(const uint32_t* a, uint32_t* r) {
    fp2 o = fp2_zero();
    bool ok = fp2_sqrt(ld<fp2>(a), o);
    st(r, o);
    return ok;
}
int hs_fp2_sgn0(const uint32_t* a) { return (int)fp2_sgn0(ld<fp2>(a)); }

// ---- Fp12
void hs_fp12_mul(const uint32_t* a, const uint32_t* b, uint32_t* r) { st(r, fp12_mul(ld<fp12>(a), ld<fp12>(b))); }
void hs_fp12_sqr(const uint32_t* a, uint32_t* r) { st(r, fp12_sqr(ld<fp12>(a))); }
void hs_fp12_inv(const uint32_t* a, uint32_t* r) { st(r, fp12_inv(ld<fp12>(a))); }
void hs_fp12_frob(const uint32_t* a, uint32_t* r) { st(r, fp12_frob(ld<fp12>(a))); }
void hs_fp12_frob2(const uint32_t* a, uint32_t* r) { st(r, fp12_frob2(ld<fp12>(a))); }
void hs_fp12_cyc_sqr(const uint32_t* a, uint32_t* r) { st(r, fp12_cyclotomic_sqr(ld<fp12>(a))); }
void hs_fp12_mul_by_014(const uint32_t* a, const uint32_t* l0, const uint32_t* l1, const uint32_t* l4, uint32_t* r) {
    st(r, fp12_mul_by_014(ld<fp12>(a), ld<fp2>(l0), ld<fp2>(l1), ld<fp2>(l4)));
}

// ---- G1 / G2 on ZCash encodings
int hs_g1_decompress(const uint8_t* in, uint32_t* aff_out) {
    g1_aff a;
    a.x = fp_zero();
    a.y = fp_zero();
    int s = g1_decompress(in, a);
    st(aff_out, a);
    return s;
}
int hs_g2_decompress(const uint8_t* in, uint32_t* aff_out) {
    g2_aff a;
    a.x = fp2_zero();
    a.y = fp2_zero();
    int s = g2_decompress(in, a);
    st(aff_out, a);
    return s;
}
// out = compress(decompress(a) + decompress(b)); returns 0, or 1 when an input is undecodable
int hs_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out, int mixed) {
    g1_aff pa, pb;
    int sa = g1_decompress(a, pa), sb = g1_decompress(b, pb);
    if (sa == DEC_BAD || sb == DEC_BAD) return 1;
    g1_jac ja = sa == DEC_INF ? pt_inf<fp>() : pt_from_affine(pa);
    g1_jac jb = sb == DEC_INF ? pt_inf<fp>() : pt_from_affine(pb);
    g1_jac r = (mixed && sb != DEC_INF) ? pt_add_mixed(ja, pb) : pt_add(ja, jb);
    g1_compress(r, out);
    return 0;
}
int hs_g2_add(const uint8_t* a, const uint8_t* b, uint8_t* out, int mixed) {
    g2_aff pa, pb;
    int sa = g2_decompress(a, pa), sb = g2_decompress(b, pb);
    if (sa == DEC_BAD || sb == DEC_BAD) return 1;
    g2_jac ja = sa == DEC_INF ? pt_inf<fp2>() : pt_from_affine(pa);
    g2_jac jb = sb == DEC_INF ? pt_inf<fp2>() : pt_from_affine(pb);
    g2_jac r = (mixed && sb != DEC_INF) ? pt_add_mixed(ja, pb) : pt_add(ja, jb);
    g2_compress(r, out);
    return 0;
}
int hs_g1_mul(const uint8_t* a, const uint32_t* k8, uint8_t* out) {
    g1_aff pa;
    int sa = g1_decompress(a, pa);
    if (sa == DEC_BAD) return 1;
    g1_jac ja = sa == DEC_INF ? pt_inf<fp>() : pt_from_affine(pa);
    g1_compress(pt_mul_var(ja, k8), out);
    return 0;
}
int hs_g2_mul(const uint8_t* a, const uint32_t* k8, uint8_t* out) {
    g2_aff pa;
    int sa = g2_decompress(a, pa);
    if (sa == DEC_BAD) return 1;
    g2_jac ja = sa == DEC_INF ? pt_inf<fp2>() : pt_from_affine(pa);
    g2_compress(pt_mul_var(ja, k8), out);
    return 0;
}
// affine Montgomery coordinates in (so that points outside the subgroup can be fed)
int hs_g1_in_subgroup_exact(const uint32_t* aff_in) { return pt_in_subgroup_exact(pt_from_affine(ld<g1_aff>(aff_in))); }
int hs_g2_in_subgroup_exact(const uint32_t* aff_in) { return pt_in_subgroup_exact(pt_from_affine(ld<g2_aff>(aff_in))); }
int hs_g2_in_subgroup_psi(const uint32_t* aff_in) { return g2_in_subgroup(pt_from_affine(ld<g2_aff>(aff_in))); }
void hs_g2_clear_cofactor(const uint32_t* aff_in, uint8_t* out96) { g2_compress(g2_clear_cofactor(pt_from_affine(ld<g2_aff>(aff_in))), out96); }

// ---- hash to curve
void hs_expand_message_xmd_256(const uint8_t* msg, uint32_t n, const uint8_t* dst, uint32_t dn, uint8_t* out) {
    expand_message_xmd_256(msg, n, dst, dn, out);
}
void hs_sha256(const uint8_t* msg, uint32_t n, uint8_t* out) {
    sha256_ctx c;
    sha256_init(c);
    sha256_update(c, msg, n);
    sha256_final(c, out);
}
void hs_hash_to_g2(const uint8_t* msg, uint32_t n, const uint8_t* dst, uint32_t dn, uint8_t* out96) {
    g2_compress(hash_to_g2(msg, n, dst, dn), out96);
}
void hs_sswu_map(const uint32_t* u, uint32_t* xy) {
    fp2 x, y;
    sswu_map(ld<fp2>(u), x, y);
    st(xy, x);
    st(xy + 24, y);
}

// ---- pairing: P = compressed G1, Q = compressed G2 -> Fp12 (Montgomery) after the final exponentiation
int hs_pairing(const uint8_t* p48, const uint8_t* q96, uint32_t* out, int do_final_exp, int scale_p) {
    g1_aff pa;
    g2_aff qa;
    int sp = g1_decompress(p48, pa), sq = g2_decompress(q96, qa);
    if (sp == DEC_BAD || sq == DEC_BAD) return 1;
    g1_jac pj = sp == DEC_INF ? pt_inf<fp>() : pt_from_affine(pa);
    if (scale_p && sp == DEC_OK) pj = pt_add_mixed(pt_dbl(pj), pa);      // 3P with Z != 1: caller compares with e(3P, Q)
    fp12 f = miller_loop(pj, qa, sq == DEC_INF);
    if (do_final_exp) f = final_exponentiation(f);
    st(out, f);
    return 0;
}
void hs_final_exp(const uint32_t* in, uint32_t* out) { st(out, final_exponentiation(ld<fp12>(in))); }

// ---- kernel cores (the per-thread bodies of the CUDA kernels), looped on the host
void hs_core_registry_load(const uint8_t* pk48, uint64_t n, uint32_t* records, uint8_t* valid) {
    for (uint64_t i = 0; i < n; i++) core_registry_load(pk48, records, valid, i);
}
void hs_core_g1_aggregate(const uint32_t* records, const uint8_t* valid, const uint32_t* members, const uint32_t* off,
                          const uint8_t* bits, uint32_t bits_stride, uint32_t n_agg, uint32_t* out_jac, uint8_t* status, uint8_t* out48) {
    for (uint32_t a = 0; a < n_agg; a++) {
        g1_jac acc = pt_inf<fp>();
        uint32_t st_ = 0, cnt = 0;
        for (uint32_t j = 0; j < off[a + 1] - off[a]; j++) core_g1_accumulate(records, valid, members, off, bits, bits_stride, a, j, acc, st_, cnt);
        core_g1_finish(acc, st_, cnt, a, out_jac, status);
        if (out48) g1_compress(acc, out48 + 48 * a);
    }
}
void hs_core_g2_aggregate(const uint8_t* sig96, const uint32_t* seg_off, uint32_t n_seg, uint8_t* out96, int32_t* seg_status) {
    for (uint32_t s = 0; s < n_seg; s++) {
        g2_jac acc = pt_inf<fp2>();
        uint32_t bad = 0;
        for (uint32_t j = seg_off[s]; j < seg_off[s + 1]; j++) core_g2_accumulate(sig96, j, acc, bad);
        core_g2_agg_finish(acc, bad, seg_off[s + 1] - seg_off[s], s, out96, seg_status);
    }
}
void hs_core_verify(const uint32_t* pk_jac, const uint8_t* pk_status, const uint8_t* msg32, const uint8_t* sig96, uint32_t n_agg, uint8_t* ok) {
    for (uint32_t a = 0; a < n_agg; a++) {
        g2_aff h, s;
        uint8_t hflag, sflag;
        core_hash_msg(msg32, a, h, hflag);
        core_sig_prepare(sig96, a, s, sflag);
        fp12 f0 = core_miller_pk(pk_jac, pk_status, a, h, hflag);
        fp12 f1 = core_miller_sig(s, sflag);
        ok[a] = core_final_verdict(f0, f1, pk_status[a], sflag);
    }
}

}  // extern "C"

extern "C" void hs_fp_sqr(const uint32_t* a, uint32_t* r) { st(r, fp_sqr(ld<fp>(a))); }

// ---- team (3-lane cooperative) pairing: three host threads + a barrier stand in for three SIMT lanes
#include <pthread.h>

#include <thread>

#include "team.cuh"
namespace b2 {
void host_team_barrier(void* b) { pthread_barrier_wait((pthread_barrier_t*)b); }
}
extern "C" int hs_team_pairing(const uint8_t* p48, const uint8_t* q96, uint32_t* out, int do_final_exp, int scale_p) {
    g1_aff pa;
    g2_aff qa;
    int sp = g1_decompress(p48, pa), sq = g2_decompress(q96, qa);
    if (sp == DEC_BAD || sq == DEC_BAD) return 1;
    g1_jac pj = sp == DEC_INF ? pt_inf<fp>() : pt_from_affine(pa);
    if (scale_p && sp == DEC_OK) pj = pt_add_mixed(pt_dbl(pj), pa);
    team_ws* ws = new team_ws();
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, 3);
    std::thread th[3];
    for (int l = 0; l < 3; l++)
        th[l] = std::thread([&, l]() {
            team tm = {l, 0u, &bar};
            team_miller_loop(tm, ws, pj, qa, sq == DEC_INF);
            if (do_final_exp) team_final_exponentiation(tm, ws);
        });
    for (int l = 0; l < 3; l++) th[l].join();
    st(out, ws->f);
    pthread_barrier_destroy(&bar);
    delete ws;
    return 0;
}
extern "C" void hs_team_final_exp(const uint32_t* in, uint32_t* out) {
    team_ws* ws = new team_ws();
    ws->f = ld<fp12>(in);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, nullptr, 3);
    std::thread th[3];
    for (int l = 0; l < 3; l++)
        th[l] = std::thread([&, l]() {
            team tm = {l, 0u, &bar};
            team_final_exponentiation(tm, ws);
        });
    for (int l = 0; l < 3; l++) th[l].join();
    st(out, ws->f);
    pthread_barrier_destroy(&bar);
    delete ws;
}

extern "C" void hs_signing_root(const uint8_t* data128, const uint8_t* domain32, uint8_t* out32) { attestation_signing_root(data128, domain32, out32); }

// ---- random-linear-combination batch verification: one group, the same per-thread bodies the kernels of csrc/rlc.cuh run
extern "C" int hs_rlc_group(const uint8_t* seed32, const uint8_t* pk48, const uint8_t* msg32, const uint8_t* sig96, uint32_t n, uint64_t* r_out,
                            uint8_t* in_batch_out, uint8_t* gpass) {
    if (n == 0 || n > 32) return 1;
    fp12* f_pk = new fp12[n];
    g2_jac ssum = pt_inf<fp2>();
    for (uint32_t i = 0; i < n; i++) {
        g1_aff pa;
        pa.x = fp_zero();
        pa.y = fp_zero();
        int sp = g1_decompress(pk48 + 48 * i, pa);
        uint8_t pk_status = (sp == DEC_OK && pt_in_subgroup_exact(pt_from_affine(pa))) ? PK_OK : PK_INVALID_KEY;
        g2_aff h, s;
        uint8_t hflag, sflag;
        core_hash_msg(msg32, i, h, hflag);
        core_sig_prepare(sig96, i, s, sflag);
        const uint64_t r = core_rlc_scalar(seed32, i, msg32 + 32 * i);
        r_out[i] = r;
        in_batch_out[i] = rlc_in_batch(pk_status, sflag) ? 1 : 0;
        f_pk[i] = fp12_one();
        if (pk_status == PK_OK) f_pk[i] = miller_loop(pt_mul_var64(pt_from_affine(pa), r), h, hflag != 0);
        if (pk_status == PK_OK && sflag == SIG_OK) ssum = pt_add(ssum, pt_mul_var64_aff(s, r));
    }
    g2_aff sa;
    sa.x = fp2_zero();
    sa.y = fp2_zero();
    const bool finite = pt_to_affine(ssum, sa);
    fp12 f_sig = core_miller_sig(sa, finite ? SIG_OK : SIG_INFINITY);
    *gpass = core_rlc_group_verdict(f_sig, f_pk, in_batch_out, n);
    delete[] f_pk;
    return 0;
}
