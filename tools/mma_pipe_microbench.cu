// mma_pipe_microbench.cu -- VERDICT round 1, "next" item 8 (exploratory, labelled as such): two thirds of a Montgomery multiplication are
// products by the CONSTANTS p' and p, i.e. in principle a batched [elements x 48 B] x [48 x 96] constant-matrix product the tensor cores
// could carry while the IMAD pipe does the variable product.  First question: do integer tensor-core MMAs and IMAD.WIDE issue concurrently
// on sm_100a at all?  (DFMA does not: tools/pipe_microbench.cu.)  Modes: IMAD.WIDE chains only; mma.sync.m16n8k32.s8 only; both in the
// same warp; alternating warps.  Reports SM clocks per mode and per-SM warp-instruction rates.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/mma_pipe_microbench.bin tools/mma_pipe_microbench.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 8192

__device__ __forceinline__ void wide_step(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}
__device__ __forceinline__ void mma_s8(int (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k32.row.col.s32.s8.s8.s32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+r"(d[0]), "+r"(d[1]), "+r"(d[2]), "+r"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int MODE>
__global__ void __launch_bounds__(128) k(uint32_t* out, uint32_t seed, long long* blk_clk) {
    const long long c0 = clock64();
    uint32_t lo[8], hi[8];
    int acc[4][4];
    uint32_t A[4] = {seed, seed * 3, seed * 5, seed * 7}, B[2] = {seed * 11, seed * 13};
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) {
        lo[k2] = k2 + a;
        hi[k2] = k2 * b;
    }
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int r = 0; r < 4; r++) acc[q][r] = q + r;
    const bool warp_is_int = ((threadIdx.x >> 5) & 1) == 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2 || (MODE == 3 && warp_is_int)) {
#pragma unroll
            for (int k2 = 0; k2 < 8; k2++) wide_step(lo[k2], hi[k2], a, b);
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && !warp_is_int)) {
#pragma unroll
            for (int q = 0; q < 4; q++) mma_s8(acc[q], A, B);          // four independent accumulators
        }
    }
    uint32_t x = 0;
#pragma unroll
    for (int k2 = 0; k2 < 8; k2++) x ^= lo[k2] ^ hi[k2];
#pragma unroll
    for (int q = 0; q < 4; q++)
#pragma unroll
        for (int r = 0; r < 4; r++) x ^= (uint32_t)acc[q][r];
    out[blockIdx.x * blockDim.x + threadIdx.x] = x;
    __syncthreads();
    if (threadIdx.x == 0) blk_clk[blockIdx.x] = clock64() - c0;
}

template <int MODE> static void run(const char* name, double wide, double mma, int n_sm, uint32_t* d_out, long long* d_clk) {
    const int blocks = n_sm * 4;
    static long long h[4096];
    for (int w = 0; w < 2; w++) k<MODE><<<blocks, 128>>>(d_out, 7, d_clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 128>>>(d_out, 7, d_clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h, d_clk, blocks * 8, cudaMemcpyDeviceToHost);
    double clks = 0;
    for (int b = 0; b < blocks; b++) clks += (double)h[b] / blocks;
    const double warps = blocks * 4.0;
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"sm_clks\": %.0f, \"wide_per_clk_sm\": %.3f, \"mma_m16n8k32_s8_per_clk_sm\": %.3f, \"cuda\": \"%s\"}\n", name, ms, clks,
           wide * ITERS * warps / n_sm / clks, mma * ITERS * warps / n_sm / clks, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    uint32_t* d_out;
    long long* d_clk;
    cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 4 * 128 * 4);
    cudaMalloc(&d_clk, 4096 * 8);
    run<0>("imad_wide_only", 8, 0, p.multiProcessorCount, d_out, d_clk);
    run<1>("mma_s8_only", 0, 4, p.multiProcessorCount, d_out, d_clk);
    run<2>("same_warp_wide+mma", 8, 4, p.multiProcessorCount, d_out, d_clk);
    run<3>("split_warps_wide|mma", 4, 2, p.multiProcessorCount, d_out, d_clk);
    return 0;
}
