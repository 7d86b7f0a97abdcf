// rlc.cuh -- kernels of the random-linear-combination batch mode of FastAggregateVerify (b2_set_verify_mode(ctx, 1, seed)).
// Per-thread bodies and the argument for the construction: cores.cuh ("random-linear-combination batch verification").
// Reference call site: bls.FastAggregateVerify inside is_valid_indexed_attestation (/root/reference/pos-evolution.md:736, :976),
// once per aggregate there; the batch form gives the same verdict vector with one final exponentiation per B2_RLC_GROUP
// aggregates (per-aggregate fallback for a group whose equation fails).
#pragma once
#include "kernels.cuh"

namespace b2 {

#define B2_RLC_GROUP 32            // aggregates per batch equation = lanes of the warp that sums their [r_i] S_i

// side stream, after K2: r_i and [r_i] PK_i (Jacobian, same layout as the K2 output)
__global__ void __launch_bounds__(128, B2_TAIL_MIN_BLOCKS) k_rlc_pk(const uint8_t* __restrict__ seed32, const uint8_t* __restrict__ msg32,
                                                                    const uint32_t* __restrict__ pk_jac, const uint8_t* __restrict__ pk_status, uint32_t n_agg,
                                                                    unsigned long long* rscal, uint32_t* pk_jac_r) {
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_agg) return;
    const uint64_t r = core_rlc_scalar(seed32, a, msg32 + 32 * (uint64_t)a);
    rscal[a] = r;
    g1_jac p = pt_inf<fp>();
    if (pk_status[a] == PK_OK) p = pt_mul_var64(load_g1_jac(pk_jac + 36 * (uint64_t)a), r);
    store_g1_jac(pk_jac_r + 36 * (uint64_t)a, p);
}

// tail stream, after the aggregate signatures are affine + subgroup-checked: one warp per group, lane = aggregate:
// [r_i] S_i, warp-shuffle tree sum, lane 0 converts the group sum to affine (the Q argument of the group's Miller loop)
__global__ void __launch_bounds__(32) k_rlc_sig(const uint32_t* __restrict__ s_aff, const uint8_t* __restrict__ sflag, const uint8_t* __restrict__ pk_status,
                                                const unsigned long long* __restrict__ rscal, uint32_t n_agg, uint32_t* sg_aff, uint8_t* sg_flag) {
    const uint32_t g = blockIdx.x, lane = threadIdx.x, a = g * B2_RLC_GROUP + lane;
    g2_jac acc = pt_inf<fp2>();
    if (a < n_agg && pk_status[a] == PK_OK && sflag[a] == SIG_OK) acc = pt_mul_var64_aff(load_g2_aff(s_aff + 48 * (uint64_t)a), rscal[a]);
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        g2_jac other = shfl_down_pod(acc, delta);
        if ((int)lane < delta) acc = pt_add(acc, other);
    }
    if (lane == 0) {
        g2_aff out;
        out.x = fp2_zero();
        out.y = fp2_zero();
        const bool finite = pt_to_affine(acc, out);
        uint32_t* o = sg_aff + 48 * (uint64_t)g;
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&out);
#pragma unroll 1
        for (int k = 0; k < 48; k++) o[k] = w[k];
        sg_flag[g] = finite ? SIG_OK : SIG_INFINITY;
    }
}

// one thread per group: product of the members' Miller values with the group's signature-side value, ONE final exponentiation
__global__ void __launch_bounds__(128, B2_TAIL_MIN_BLOCKS) k_rlc_final(const uint32_t* __restrict__ f, const uint32_t* __restrict__ f_g,
                                                                       const uint8_t* __restrict__ pk_status, const uint8_t* __restrict__ sflag, uint32_t n_agg,
                                                                       uint32_t n_groups, uint8_t* gpass) {
    const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= n_groups) return;
    fp12 F;
    {
        uint32_t* w = reinterpret_cast<uint32_t*>(&F);
        const uint32_t* p = f_g + 144 * (2 * (uint64_t)g + 1);
#pragma unroll 1
        for (int k = 0; k < 144; k++) w[k] = p[k];
    }
    bool any = false;
#pragma unroll 1
    for (uint32_t a = g * B2_RLC_GROUP; a < min(n_agg, (g + 1) * B2_RLC_GROUP); a++) {
        if (!rlc_in_batch(pk_status[a], sflag[a])) continue;
        fp12 m;
        uint32_t* w = reinterpret_cast<uint32_t*>(&m);
        const uint32_t* p = f + 144 * (2 * (uint64_t)a);
#pragma unroll 1
        for (int k = 0; k < 144; k++) w[k] = p[k];
        F = fp12_mul(F, m);
        any = true;
    }
    gpass[g] = !any ? 1 : (fp12_is_one(final_exponentiation(F)) ? 1 : 0);
}

// verdicts of the groups that passed (and of the aggregates rejected outright); the members of a failed group are decided by
// the masked per-aggregate kernels that follow
__global__ void __launch_bounds__(128) k_rlc_verdict(const uint8_t* __restrict__ pk_status, const uint8_t* __restrict__ sflag,
                                                     const uint8_t* __restrict__ gpass, uint32_t n_agg, uint8_t* ok) {
    uint32_t a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= n_agg) return;
    ok[a] = (rlc_in_batch(pk_status[a], sflag[a]) && gpass[a / B2_RLC_GROUP]) ? 1 : 0;
}

}  // namespace b2
