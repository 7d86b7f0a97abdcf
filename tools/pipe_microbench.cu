// pipe_microbench.cu -- which sm_100a pipes can the 381-bit field arithmetic use at the same time?
// Measures warp-instruction throughput per SM per clock of (A) the IMAD.WIDE.U32 carry chain fp.cuh is made of, (B) DFMA,
// (C) both in the same thread, (D) both in different warps of one SM, (E) IADD3 carry chains beside IMAD.WIDE.
// Build:  nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o pipe_microbench tools/pipe_microbench.cu ; run on the B200.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 16384

__device__ __forceinline__ void wide_step(uint32_t& lo, uint32_t& hi, uint32_t a, uint32_t b) {
    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.u32 %1, %2, %3, %1;" : "+r"(lo), "+r"(hi) : "r"(a), "r"(b));
}

template <int MODE>
__global__ void __launch_bounds__(128) k_pipes(uint32_t* out, uint32_t seed, double dseed, long long* blk_clk) {
    const long long c0 = clock64();
    uint32_t lo[8], hi[8];
    double d[16];
    uint32_t s[8];
    uint32_t a = seed + threadIdx.x, b = seed * 3 + 1;
    double da = dseed + threadIdx.x, db = dseed * 0.5;
#pragma unroll
    for (int k = 0; k < 8; k++) {
        lo[k] = k + a;
        hi[k] = k * b;
        s[k] = k ^ a;
    }
#pragma unroll
    for (int k = 0; k < 16; k++) d[k] = k * dseed;
    const bool warp_is_int = ((threadIdx.x >> 5) & 1) == 0;
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0 || MODE == 2 || MODE == 4 || (MODE == 3 && warp_is_int)) {
#pragma unroll
            for (int k = 0; k < 8; k++) wide_step(lo[k], hi[k], a, b);
        }
        if (MODE == 1 || MODE == 2 || (MODE == 3 && !warp_is_int)) {
#pragma unroll
            for (int k = 0; k < 16; k++) d[k] = fma(d[k], da, db);
        }
        if (MODE == 4 || MODE == 5) {
#pragma unroll
            for (int k = 0; k < 8; k++) {
                asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %0, %0, %2;" : "+r"(s[k]) : "r"(a), "r"(b));
                asm volatile("add.cc.u32 %0, %0, %1;\n\taddc.u32 %0, %0, %2;" : "+r"(s[k]) : "r"(b), "r"(a));
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) acc ^= lo[k] ^ hi[k] ^ s[k];
    double dacc = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) dacc += d[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc ^ (uint32_t)__double_as_longlong(dacc);
    __syncthreads();
    if (threadIdx.x == 0) blk_clk[blockIdx.x] = clock64() - c0;
}

// dependent-issue latency: one warp per SM, one chain
__global__ void k_lat_dfma(double* out, double da, double db, long long* clk) {
    double d = da;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) d = fma(d, da, db);
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = d;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
__global__ void k_lat_wide(uint32_t* out, uint32_t a, uint32_t b, long long* clk) {
    uint32_t lo = a, hi = b;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < 1024; it++) {
#pragma unroll
        for (int k = 0; k < 16; k++) wide_step(lo, hi, a, b);
    }
    long long t1 = clock64();
    out[blockIdx.x * 32 + threadIdx.x] = lo ^ hi;
    if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}

template <int MODE>
static void run(const char* name, double wide_per_iter, double dfma_per_iter, double iadd_per_iter, int n_sm, double clk_ghz, uint32_t* d_out) {
    int blocks = n_sm * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    static long long* d_clk = nullptr;
    static long long h_clk[4096];
    if (!d_clk) cudaMalloc(&d_clk, sizeof(h_clk));
    for (int w = 0; w < 3; w++) k_pipes<MODE><<<blocks, 128>>>(d_out, 7, 1.000001, d_clk);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    k_pipes<MODE><<<blocks, 128>>>(d_out, 7, 1.000001, d_clk);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h_clk, d_clk, blocks * 8, cudaMemcpyDeviceToHost);
    double warps = blocks * 4.0;
    double clks = 0;                          // SM clocks the resident blocks were alive for (frequency-independent)
    for (int b = 0; b < blocks; b++) clks += (double)h_clk[b] / blocks;
    (void)clk_ghz;
    // per-SM warp-instructions per clock
    printf("{\"mode\": \"%s\", \"ms\": %.4f, \"sm_clks\": %.0f, \"wide_per_clk_sm\": %.3f, \"dfma_per_clk_sm\": %.3f, \"iadd_per_clk_sm\": %.3f}\n", name, ms, clks,
           wide_per_iter * ITERS * warps / n_sm / clks, dfma_per_iter * ITERS * warps / n_sm / clks, iadd_per_iter * ITERS * warps / n_sm / clks);
}

int main() {
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    int clk_khz = 0;
    cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double ghz = clk_khz * 1e-6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    uint32_t* d_out;
    cudaMalloc(&d_out, (size_t)p.multiProcessorCount * 4 * 128 * 4);
    // "wide" counts IMAD.WIDE pairs as ONE instruction each when ptxas fuses them (check with cuobjdump)
    run<0>("imad_wide_only", 8, 0, 0, p.multiProcessorCount, ghz, d_out);
    run<1>("dfma_only", 0, 16, 0, p.multiProcessorCount, ghz, d_out);
    run<2>("same_thread_wide+dfma", 8, 16, 0, p.multiProcessorCount, ghz, d_out);
    run<3>("split_warps_wide|dfma", 4, 8, 0, p.multiProcessorCount, ghz, d_out);
    run<4>("same_thread_wide+iadd", 8, 0, 32, p.multiProcessorCount, ghz, d_out);
    run<5>("iadd_only", 0, 0, 32, p.multiProcessorCount, ghz, d_out);
    long long* d_clk;
    long long h_clk = 0;
    cudaMalloc(&d_clk, 8);
    k_lat_dfma<<<1, 32>>>((double*)d_out, 1.0000001, 0.5, d_clk);
    cudaMemcpy(&h_clk, d_clk, 8, cudaMemcpyDeviceToHost);
    printf("{\"mode\": \"dfma_dependent_latency_clk\", \"clk\": %.2f}\n", h_clk / (1024.0 * 16));
    k_lat_wide<<<1, 32>>>(d_out, 12345, 678, d_clk);
    cudaMemcpy(&h_clk, d_clk, 8, cudaMemcpyDeviceToHost);
    printf("{\"mode\": \"imad_wide_dependent_latency_clk\", \"clk\": %.2f}\n", h_clk / (1024.0 * 16));
    return 0;
}
