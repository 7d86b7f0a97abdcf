// gather_bench.cu -- microbenchmark of the pubkey-gather stage alone (north_star: ">= 60 % of the HBM-read roofline on the pubkey
// gather"): 2^20 records of 96 B at the positions of a random permutation, 2 048 rows of 512, every record read exactly once
// per launch, XOR-checksummed per row.  Which way of issuing a random 96-byte gather gets closest to the measured HBM peak?
//
//   0  lane-per-record, one warp per row, six LDG.E.128 per lane              (the form of k_g1_gather_ldg_probe)
//   1  lane-per-record, four warps per row, 4 records per lane all in flight
//   2  chunk-per-lane: 6 consecutive lanes read the six 16-byte chunks of one record (coalesced within a record), 4 warps per row
//   3  cp.async.bulk per record, one warp per row, 4 stages                     (the form of k_g1_gather_tma)
//   4  cp.async.bulk per record, four warps per row, the whole row in flight at once
//   5  cp.async.bulk.tensor tile::gather4 (4 records per TMA instruction), four warps per row, the whole row in flight
//
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/gather_bench.bin tools/gather_bench.cu -lcuda
// Run:    tools/gather_bench.bin            (prints one JSON line per variant; L2 flushed before every timed launch)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define CKC(x)                                                                                  \
    do {                                                                                        \
        cudaError_t e_ = (x);                                                                   \
        if (e_ != cudaSuccess) {                                                                \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__);     \
            exit(1);                                                                            \
        }                                                                                       \
    } while (0)

static constexpr uint32_t N = 1u << 20, ROW = 512, NROWS = N / ROW;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    asm volatile(
        "{\n.reg .pred p;\nW_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D_%=;\nbra W_%=;\nD_%=:\n}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_bulk(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes),
                 "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ uint32_t xor4(uint4 v) { return v.x ^ v.y ^ v.z ^ v.w; }
__device__ __forceinline__ uint32_t warp_xor(uint32_t v) {
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) v ^= __shfl_down_sync(0xffffffffu, v, d);
    return v;
}

// ---- 0: lane per record, one warp per row
__global__ void __launch_bounds__(128) g0(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, uint32_t* out) {
    const int lane = threadIdx.x & 31, row = blockIdx.x * 4 + (threadIdx.x >> 5);
    uint32_t chk = 0;
#pragma unroll 2
    for (int j = lane; j < ROW; j += 32) {
        const uint4* r = rec + 6ull * idx[row * ROW + j];
#pragma unroll
        for (int k = 0; k < 6; k++) chk ^= xor4(__ldg(r + k));
    }
    chk = warp_xor(chk);
    if (lane == 0) out[row] = chk;
}
// ---- 1: lane per record, block of 128 per row, 4 records per lane, all 24 loads issued before use
__global__ void __launch_bounds__(128) g1(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, uint32_t* out) {
    __shared__ uint32_t part[4];
    const int row = blockIdx.x, t = threadIdx.x;
    uint32_t id[4];
#pragma unroll
    for (int q = 0; q < 4; q++) id[q] = idx[row * ROW + q * 128 + t];
    uint4 v[4][6];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 6; k++) v[q][k] = __ldg(rec + 6ull * id[q] + k);
    uint32_t chk = 0;
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int k = 0; k < 6; k++) chk ^= xor4(v[q][k]);
    chk = warp_xor(chk);
    if ((t & 31) == 0) part[t >> 5] = chk;
    __syncthreads();
    if (t == 0) out[row] = part[0] ^ part[1] ^ part[2] ^ part[3];
}
// ---- 2: chunk per lane: thread t of the block reads 16-byte chunk (t % 6) of record (t / 6) of each group of 21 records (126 lanes)
__global__ void __launch_bounds__(128) g2(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, uint32_t* out) {
    __shared__ uint32_t part[4];
    __shared__ uint32_t sidx[ROW];
    const int row = blockIdx.x, t = threadIdx.x;
#pragma unroll
    for (int q = 0; q < 4; q++) sidx[q * 128 + t] = idx[row * ROW + q * 128 + t];
    __syncthreads();
    // 512 records * 6 chunks = 3072 chunk loads = 24 per thread, chunk c of the row -> record c / 6, part c % 6
    uint4 v[24];
#pragma unroll
    for (int q = 0; q < 24; q++) {
        const int c = q * 128 + t;
        v[q] = __ldg(rec + 6ull * sidx[c / 6] + (c % 6));
    }
    uint32_t chk = 0;
#pragma unroll
    for (int q = 0; q < 24; q++) chk ^= xor4(v[q]);
    chk = warp_xor(chk);
    if ((t & 31) == 0) part[t >> 5] = chk;
    __syncthreads();
    if (t == 0) out[row] = part[0] ^ part[1] ^ part[2] ^ part[3];
}
// ---- 3: bulk copy per record, one warp per row, NST stages of 32 records
#define NST 4
#define SLOT 112
struct ws3 {
    unsigned long long bar[NST];
    uint8_t rec[NST][32 * SLOT];
};
__global__ void __launch_bounds__(128) g3(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, row = blockIdx.x * 4 + warp;
    ws3* ws = reinterpret_cast<ws3*>(sm) + warp;
    if (lane == 0) {
        for (int s = 0; s < NST; s++) mbar_init(&ws->bar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    uint32_t chk = 0, ph = 0;
    auto issue = [&](int it) {
        const int st = it % NST;
        if (lane == 0) mbar_expect_tx(&ws->bar[st], 32 * 96);
        tma_bulk(ws->rec[st] + lane * SLOT, rec + 6ull * idx[row * ROW + it * 32 + lane], 96, &ws->bar[st]);
    };
    for (int it = 0; it < NST - 1; it++) issue(it);
    for (int it = 0; it < ROW / 32; it++) {
        if (it + NST - 1 < ROW / 32) issue(it + NST - 1);
        const int st = it % NST;
        mbar_wait(&ws->bar[st], (ph >> st) & 1);
        ph ^= 1u << st;
        const uint4* r = reinterpret_cast<const uint4*>(ws->rec[st] + lane * SLOT);
#pragma unroll
        for (int k = 0; k < 6; k++) chk ^= xor4(r[k]);
        __syncwarp();
    }
    chk = warp_xor(chk);
    if (lane == 0) out[row] = chk;
}
// ---- 4: bulk copy per record, block of 128 per row, the whole row (512 x 96 B) in flight at once, one mbarrier
__global__ void __launch_bounds__(128) g4(const uint4* __restrict__ rec, const uint32_t* __restrict__ idx, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t part[4];
    const int row = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) mbar_expect_tx(&bar, ROW * 96);
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const int j = q * 128 + t;
        tma_bulk(sm + j * SLOT, rec + 6ull * idx[row * ROW + j], 96, &bar);
    }
    mbar_wait(&bar, 0);
    uint32_t chk = 0;
#pragma unroll
    for (int q = 0; q < 4; q++) {
        const uint4* r = reinterpret_cast<const uint4*>(sm + (q * 128 + t) * SLOT);
#pragma unroll
        for (int k = 0; k < 6; k++) chk ^= xor4(r[k]);
    }
    chk = warp_xor(chk);
    if ((t & 31) == 0) part[t >> 5] = chk;
    __syncthreads();
    if (t == 0) out[row] = part[0] ^ part[1] ^ part[2] ^ part[3];
}
// ---- 5: tile::gather4: one TMA instruction fetches 4 rows of the [N][24 x u32] tensor; block of 128 per row, thread t issues record
// group t (4 records), whole row in flight; destination = 4 x 96 B contiguous per instruction
__global__ void __launch_bounds__(128) g5(const __grid_constant__ CUtensorMap tmap, const uint32_t* __restrict__ idx, uint32_t* out) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ unsigned long long bar;
    __shared__ uint32_t part[4];
    const int row = blockIdx.x, t = threadIdx.x;
    if (t == 0) {
        mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) mbar_expect_tx(&bar, ROW * 96);
    const uint4 id = reinterpret_cast<const uint4*>(idx + row * ROW)[t];
    asm volatile("cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
                     smem_u32(sm + t * 384)),
                 "l"(&tmap), "r"(0), "r"((int)id.x), "r"((int)id.y), "r"((int)id.z), "r"((int)id.w), "r"(smem_u32(&bar))
                 : "memory");
    mbar_wait(&bar, 0);
    uint32_t chk = 0;
    const uint4* r = reinterpret_cast<const uint4*>(sm + t * 384);
#pragma unroll
    for (int k = 0; k < 24; k++) chk ^= xor4(r[k]);
    chk = warp_xor(chk);
    if ((t & 31) == 0) part[t >> 5] = chk;
    __syncthreads();
    if (t == 0) out[row] = part[0] ^ part[1] ^ part[2] ^ part[3];
}

int main() {
    CKC(cudaSetDevice(0));
    std::vector<uint32_t> h_rec((size_t)N * 24), h_idx(N), h_want(NROWS, 0);
    uint64_t s = 0x9e3779b97f4a7c15ull;
    auto rnd = [&]() {
        s ^= s << 13; s ^= s >> 7; s ^= s << 17;
        return s;
    };
    for (auto& w : h_rec) w = (uint32_t)rnd();
    for (uint32_t i = 0; i < N; i++) h_idx[i] = i;
    for (uint32_t i = N - 1; i > 0; i--) std::swap(h_idx[i], h_idx[rnd() % (i + 1)]);
    for (uint32_t r = 0; r < NROWS; r++)
        for (uint32_t j = 0; j < ROW; j++)
            for (int k = 0; k < 24; k++) h_want[r] ^= h_rec[(size_t)h_idx[r * ROW + j] * 24 + k];
    uint32_t *d_rec, *d_idx, *d_out;
    uint8_t* d_flush;
    const size_t flush_bytes = 512ull << 20;
    CKC(cudaMalloc(&d_rec, (size_t)N * 96));
    CKC(cudaMalloc(&d_idx, (size_t)N * 4));
    CKC(cudaMalloc(&d_out, NROWS * 4));
    CKC(cudaMalloc(&d_flush, flush_bytes));
    CKC(cudaMemcpy(d_rec, h_rec.data(), (size_t)N * 96, cudaMemcpyHostToDevice));
    CKC(cudaMemcpy(d_idx, h_idx.data(), (size_t)N * 4, cudaMemcpyHostToDevice));
    // tensor map for variant 5: 2-D [N rows][24 u32], box {24, 1}
    CUtensorMap tmap;
    bool have_tmap = false;
    {
        cuInit(0);
        cuuint64_t gdim[2] = {24, N};
        cuuint64_t gstride[1] = {96};
        cuuint32_t box[2] = {24, 1};
        cuuint32_t estr[2] = {1, 1};
        CUresult rc = cuTensorMapEncodeTiled(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, d_rec, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        have_tmap = rc == CUDA_SUCCESS;
        if (!have_tmap) printf("{\"variant\": 5, \"skipped\": \"cuTensorMapEncodeTiled rc=%d\"}\n", (int)rc);
    }
    CKC(cudaFuncSetAttribute(g3, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)sizeof(ws3)));
    CKC(cudaFuncSetAttribute(g4, cudaFuncAttributeMaxDynamicSharedMemorySize, ROW * SLOT));
    CKC(cudaFuncSetAttribute(g5, cudaFuncAttributeMaxDynamicSharedMemorySize, ROW * 96));
    const double bytes = (double)N * 96 + (double)N * 4;
    const char* names[6] = {"ldg lane-per-record, warp per row", "ldg lane-per-record, block per row, all loads in flight", "ldg chunk-per-lane (coalesced within a record), block per row",
                            "cp.async.bulk per record, warp per row, 4 stages", "cp.async.bulk per record, block per row, whole row in flight",
                            "cp.async.bulk.tensor tile::gather4, block per row, whole row in flight"};
    cudaEvent_t e0, e1;
    CKC(cudaEventCreate(&e0));
    CKC(cudaEventCreate(&e1));
    for (int v = 0; v < 6; v++) {
        if (v == 5 && !have_tmap) continue;
        float best = 1e9f, sum = 0;
        float warm = 1e9f;
        const int reps = 12;
        bool ok = true;
        for (int it = 0; it < 2 * reps; it++) {
            const bool cold = it < reps;
            if (cold) CKC(cudaMemsetAsync(d_flush, it, flush_bytes));
            CKC(cudaMemsetAsync(d_out, 0, NROWS * 4));
            CKC(cudaEventRecord(e0));
            switch (v) {
                case 0: g0<<<NROWS / 4, 128>>>((const uint4*)d_rec, d_idx, d_out); break;
                case 1: g1<<<NROWS, 128>>>((const uint4*)d_rec, d_idx, d_out); break;
                case 2: g2<<<NROWS, 128>>>((const uint4*)d_rec, d_idx, d_out); break;
                case 3: g3<<<NROWS / 4, 128, 4 * sizeof(ws3)>>>((const uint4*)d_rec, d_idx, d_out); break;
                case 4: g4<<<NROWS, 128, ROW * SLOT>>>((const uint4*)d_rec, d_idx, d_out); break;
                case 5: g5<<<NROWS, 128, ROW * 96>>>(tmap, d_idx, d_out); break;
            }
            CKC(cudaEventRecord(e1));
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) {
                printf("{\"variant\": %d, \"error\": \"%s\"}\n", v, cudaGetErrorString(e));
                return 0;       // a faulted context cannot run the remaining variants
            }
            float ms;
            CKC(cudaEventElapsedTime(&ms, e0, e1));
            if (cold) {
                best = std::min(best, ms);
                sum += ms;
            } else {
                warm = std::min(warm, ms);
            }
            if (it == 0) {
                std::vector<uint32_t> got(NROWS);
                CKC(cudaMemcpy(got.data(), d_out, NROWS * 4, cudaMemcpyDeviceToHost));
                ok = got == h_want;
            }
        }
        const double mean = sum / reps;
        printf("{\"variant\": %d, \"name\": \"%s\", \"checksum_ok\": %s, \"cold_us_mean\": %.2f, \"cold_us_best\": %.2f, \"cold_GBps_mean\": %.1f, \"cold_GBps_best\": %.1f, "
               "\"warm_us_best\": %.2f, \"warm_GBps_best\": %.1f}\n",
               v, names[v], ok ? "true" : "false", mean * 1e3, best * 1e3, bytes / (mean * 1e-3) / 1e9, bytes / (best * 1e-3) / 1e9, warm * 1e3,
               bytes / (warm * 1e-3) / 1e9);
        fflush(stdout);
    }
    return 0;
}
