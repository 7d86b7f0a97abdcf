// gather.cuh -- K2 with the pubkey table staged through TMA (cp.async.bulk) into shared memory.
//
// north_star: "coalesced 128-bit HBM loads of attestation aggregation_bits and the validator pubkey / effective-balance
// tables staged through TMA into shared memory", ">= 60 % of HBM-read roofline on the pubkey gather".  The reference
// side of this is get_attesting_indices + the pubkey list comprehension of is_valid_indexed_attestation
// (/root/reference/pos-evolution.md:736, :745; committees :495-504).
//
// One WARP per aggregate.  Per 512-member chunk of the committee:
//   1. the bit row (64 B) and the u32 member indices (2 KB) are brought into shared memory with 128-bit loads
//      (LDG.E.128 -> STS.128; scalar loads when a row is not 16-byte aligned);
//   2. the 96-byte affine pubkey records of the SET bits are fetched by the TMA engine: one `cp.async.bulk` (UBLKCP) per
//      record, global -> shared, completing on an mbarrier whose transaction count lane 0 arms with popc(set bits) * 96 B.
//      NST stages of 32 records are in flight per warp, so the random 96-byte reads overlap each other and the arithmetic;
//   3. lane l reads record l of the stage with six 128-bit LDS (slots padded to 112 B: conflict-free for a quarter warp) and
//      adds it to its Jacobian partial sum (mixed addition, 11 Fp multiplications);
//   4. warp-shuffle tree over the 32 partial sums.
// An invalid registry entry is stored as the all-zero record (core_registry_load), which is not a curve point, so the
// gather needs no second random read of the `valid` byte table: zero record <=> KeyValidate failed.
//
// k_g1_gather_probe is the same staging with the additions replaced by an XOR checksum: the GATHER STAGE ALONE, whose
// achieved bytes/s against the measured HBM peak is the `roofline.gather` object of bench.py.  With the additions the kernel is
// bound by the integer multiply pipe (34 MAC/B), and the two numbers are reported side by side.
#pragma once
#include <cuda_runtime.h>

#include "cores.cuh"

namespace b2 {

#define B2_GATHER_NST 4             // stages of 32 records in flight per warp
#define B2_GATHER_SLOT 112          // bytes per record slot in shared memory (96 B record + 16 B pad)
#define B2_GATHER_CHUNK 512         // members per staged chunk of indices / bits
struct gather_ws {
    unsigned long long bar[B2_GATHER_NST];                          // one mbarrier per stage
    unsigned long long pad_[2];
    uint32_t idx[B2_GATHER_CHUNK];                                  // member indices of the chunk
    uint32_t bits[B2_GATHER_CHUNK / 32];                            // bit row of the chunk, one word per 32 members
    uint8_t rec[B2_GATHER_NST][32 * B2_GATHER_SLOT];                // record slots
};
static_assert(sizeof(gather_ws) % 16 == 0, "per-warp workspaces must keep 16-byte alignment");

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// TMA bulk copy global -> shared (size multiple of 16, both addresses 16-byte aligned), completion bytes on `bar`
__device__ __forceinline__ void tma_load_bulk(void* dst_smem, const void* src_gmem, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// stage the bit row and the member indices of members [c0, c0+cn) of aggregate a (row starts: brow, mrow)
__device__ __forceinline__ void gather_stage_rows(gather_ws* ws, const uint8_t* brow, const uint32_t* mrow, uint32_t c0, uint32_t cn, int lane) {
    const uint32_t* msrc = mrow + c0;
    if ((reinterpret_cast<uintptr_t>(msrc) & 15u) == 0) {
        const uint4* s4 = reinterpret_cast<const uint4*>(msrc);
        uint4* d4 = reinterpret_cast<uint4*>(ws->idx);
        for (uint32_t k = lane; k < cn / 4; k += 32) d4[k] = __ldg(s4 + k);                  // LDG.E.128
        for (uint32_t k = (cn & ~3u) + lane; k < cn; k += 32) ws->idx[k] = __ldg(msrc + k);
    } else {
        for (uint32_t k = lane; k < cn; k += 32) ws->idx[k] = __ldg(msrc + k);
    }
    const uint8_t* bsrc = brow + (c0 >> 3);                                                  // c0 is a multiple of 512
    const uint32_t nbytes = (cn + 7) >> 3;
    if ((reinterpret_cast<uintptr_t>(bsrc) & 15u) == 0 && (nbytes & 15u) == 0) {
        if ((uint32_t)lane < nbytes / 16) reinterpret_cast<uint4*>(ws->bits)[lane] = __ldg(reinterpret_cast<const uint4*>(bsrc) + lane);
    } else {
        uint8_t* db = reinterpret_cast<uint8_t*>(ws->bits);
        for (uint32_t k = lane; k < ((nbytes + 3) & ~3u); k += 32) db[k] = k < nbytes ? __ldg(bsrc + k) : (uint8_t)0;
    }
    __syncwarp();
}

// issue the record copies of iteration `it` (members it*32 .. it*32+31 of the chunk) into stage it % NST
__device__ __forceinline__ void gather_issue(gather_ws* ws, const uint32_t* __restrict__ records, uint32_t it, uint32_t cn, uint64_t n_val, int lane,
                                             uint32_t& bad_index) {
    const uint32_t j = it * 32 + lane;
    const uint32_t st = it % B2_GATHER_NST;
    bool want = j < cn && ((ws->bits[it] >> lane) & 1u);
    uint32_t idx = 0;
    if (want) {
        idx = ws->idx[j];
        if (idx >= n_val) {
            bad_index = 1;
            want = false;
        }
    }
    const unsigned m = __ballot_sync(0xffffffffu, want);
    if (lane == 0) mbar_expect_tx(&ws->bar[st], __popc(m) * 96u);          // completes at once when no record is wanted
    if (want) tma_load_bulk(ws->rec[st] + lane * B2_GATHER_SLOT, records + 24 * (uint64_t)idx, 96u, &ws->bar[st]);
}

// PROBE = false: K2 (pubkey aggregation);  PROBE = true: gather stage alone (XOR checksum of the fetched records)
template <bool PROBE>
__global__ void __launch_bounds__(128) k_g1_gather_tma(const uint32_t* __restrict__ records, const uint32_t* __restrict__ members,
                                                        const uint32_t* __restrict__ off, const uint8_t* __restrict__ bits, uint32_t bits_stride,
                                                        uint32_t n_agg, uint32_t* out_jac, uint8_t* out_status, uint64_t n_val, uint32_t* guard,
                                                        uint32_t* probe_out) {
    extern __shared__ __align__(128) unsigned char gather_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t a = blockIdx.x * (blockDim.x >> 5) + warp;
    if (a >= n_agg) return;
    gather_ws* ws = reinterpret_cast<gather_ws*>(gather_smem) + warp;
    if (lane == 0) {
#pragma unroll
        for (int s = 0; s < B2_GATHER_NST; s++) mbar_init(&ws->bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    const uint32_t o0 = off[a], o1 = off[a + 1];
    uint32_t size = o1 - o0, status = 0, cnt = 0, bad_index = 0, chk = 0;
    if (o1 < o0 || size > bits_stride * 8u) {
        if (lane == 0 && guard) atomicOr(guard, (uint32_t)GUARD_BAD_ROW);
        size = 0;
        status = PK_BAD_INDEX;
    }
    g1_jac acc = pt_inf<fp>();
    uint32_t phase_bits = 0;                                    // bit s = parity to wait for on stage s
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < size; c0 += B2_GATHER_CHUNK) {
        const uint32_t cn = min((uint32_t)B2_GATHER_CHUNK, size - c0);
        gather_stage_rows(ws, bits + (uint64_t)a * bits_stride, members + o0, c0, cn, lane);
        const uint32_t n_it = (cn + 31) / 32;
#pragma unroll 1
        for (uint32_t it = 0; it < min(n_it, (uint32_t)B2_GATHER_NST - 1); it++) gather_issue(ws, records, it, cn, n_val, lane, bad_index);
#pragma unroll 1
        for (uint32_t it = 0; it < n_it; it++) {
            if (it + B2_GATHER_NST - 1 < n_it) gather_issue(ws, records, it + B2_GATHER_NST - 1, cn, n_val, lane, bad_index);
            const uint32_t st = it % B2_GATHER_NST;
            mbar_wait(&ws->bar[st], (phase_bits >> st) & 1u);
            phase_bits ^= 1u << st;
            const uint32_t j = it * 32 + lane;
            if (j < cn && ((ws->bits[it] >> lane) & 1u)) {
                cnt++;
                if (ws->idx[j] < n_val) {
                    const uint4* r4 = reinterpret_cast<const uint4*>(ws->rec[st] + lane * B2_GATHER_SLOT);
                    g1_aff p;
                    uint32_t* w = reinterpret_cast<uint32_t*>(&p);
                    uint32_t nz = 0;
#pragma unroll
                    for (int k = 0; k < 6; k++) {
                        const uint4 v = r4[k];                                   // LDS.128
                        w[4 * k] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
                        nz |= v.x | v.y | v.z | v.w;
                    }
                    if (PROBE) {
#pragma unroll
                        for (int k = 0; k < 24; k++) chk ^= w[k];
                    } else if (nz == 0) {
                        status |= PK_INVALID_KEY;                                // the zero record marks a key that failed KeyValidate
                    } else {
                        acc = pt_add_mixed(acc, p);
                    }
                }
            }
            __syncwarp();                                        // every lane is done with the stage before it is refilled
        }
    }
    bad_index = __reduce_or_sync(0xffffffffu, bad_index);
    if (bad_index) status |= PK_BAD_INDEX;
    if (PROBE) {
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) chk ^= __shfl_down_sync(0xffffffffu, chk, d);
        if (lane == 0) probe_out[a] = chk;
        return;
    }
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        g1_jac other = shfl_down_pod(acc, delta);
        if (lane < delta) acc = pt_add(acc, other);
    }
    status = __reduce_or_sync(0xffffffffu, status);
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    if (lane == 0) {
        if (bad_index && guard) atomicOr(guard, (uint32_t)GUARD_BAD_INDEX);
        core_g1_finish(acc, status, cnt, a, out_jac, out_status);
    }
}

// The same checksum with plain loads, in the fastest form tools/gather_bench.cu found on the B200 (profiles/r2_gather_microbench.jsonl:
// 2.83 TB/s cold against 1.58 TB/s for the TMA form above and 2.34 TB/s for a lane-per-record LDG form): one 128-thread block per
// aggregate, the u32 indices of a 512-member chunk staged in shared memory, then thread t reads 16-byte chunk (c % 6) of selected
// record (c / 6) for c = t, t + 128, ...: six consecutive lanes cover one 96-byte record, so a warp instruction touches ~11 lines
// instead of 32, and all 24 loads of a thread are in flight before the first is used.  A/B partner of the TMA probe and the
// cross-check that both fetch the same bytes.
__global__ void __launch_bounds__(128) k_g1_gather_ldg_probe(const uint32_t* __restrict__ records, const uint32_t* __restrict__ members,
                                                              const uint32_t* __restrict__ off, const uint8_t* __restrict__ bits, uint32_t bits_stride,
                                                              uint32_t n_agg, uint64_t n_val, uint32_t* probe_out) {
    __shared__ uint32_t sel[B2_GATHER_CHUNK];          // registry indices of the SET members of the chunk, compacted
    __shared__ uint32_t n_sel, part[4];
    const uint32_t a = blockIdx.x, t = threadIdx.x;
    if (a >= n_agg) return;
    const uint32_t o0 = off[a], o1 = off[a + 1];
    const uint32_t size = (o1 < o0 || o1 - o0 > bits_stride * 8u) ? 0u : o1 - o0;
    uint32_t chk = 0;
    for (uint32_t c0 = 0; c0 < size; c0 += B2_GATHER_CHUNK) {
        const uint32_t cn = min((uint32_t)B2_GATHER_CHUNK, size - c0);
        if (t == 0) n_sel = 0;
        __syncthreads();
        for (uint32_t j = t; j < cn; j += 128) {
            const uint32_t m = c0 + j;
            if ((bits[(uint64_t)a * bits_stride + (m >> 3)] >> (m & 7)) & 1) {
                const uint32_t idx = members[o0 + m];
                if (idx < n_val) sel[atomicAdd(&n_sel, 1u)] = idx;      // order is irrelevant to an XOR
            }
        }
        __syncthreads();
        const uint32_t n_chunks = n_sel * 6;
        const uint4* rec4 = reinterpret_cast<const uint4*>(records);
        uint4 v[24];
#pragma unroll
        for (int q = 0; q < 24; q++) {
            const uint32_t c = q * 128 + t;
            v[q] = c < n_chunks ? __ldg(rec4 + 6ull * sel[c / 6] + (c % 6)) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int q = 0; q < 24; q++) chk ^= v[q].x ^ v[q].y ^ v[q].z ^ v[q].w;
        __syncthreads();
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) chk ^= __shfl_down_sync(0xffffffffu, chk, d);
    if ((t & 31) == 0) part[t >> 5] = chk;
    __syncthreads();
    if (t == 0) probe_out[a] = part[0] ^ part[1] ^ part[2] ^ part[3];
}

}  // namespace b2
