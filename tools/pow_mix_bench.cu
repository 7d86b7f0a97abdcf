// pow_mix_bench.cu -- does running the Fp exponentiation chain on the FP64 pipe in half of the warps add throughput?
// Every thread computes a^((p-3)/4) for its own a; a warp takes its work in chunks of 32 from a global counter, so faster
// warps take more.  mode 0: all warps integer form (fp.cuh), 1: all warps FP64 form (fpd.cuh), 2: blocks alternate per SM.
// Also checks that both forms give identical limbs.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I pos_evolution_b200/csrc -I tools/experiments -o /tmp/pmix tools/pow_mix_bench.cu   (add -DB2_SQR_KARATSUBA=1 to time the Karatsuba squaring)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "fpd.cuh"
using namespace b2;

__global__ void __launch_bounds__(128, 4) k_pow_mix(const uint32_t* in, uint32_t* out, uint32_t n, int mode, unsigned int* counter,
                                                   unsigned int* sm_arrivals, unsigned int* done_by_form) {
    __shared__ int s_form;
    if (threadIdx.x == 0) {
        uint32_t smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        s_form = mode == 2 ? (int)(atomicAdd(&sm_arrivals[smid], 1u) & 1u) : mode;
    }
    __syncthreads();
    const int form = s_form;
    const int lane = threadIdx.x & 31;
    for (;;) {
        unsigned int base = 0;
        if (lane == 0) base = atomicAdd(counter, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        if (i < n) {
            fp a;
#pragma unroll
            for (int k = 0; k < 12; k++) a.l[k] = in[(size_t)i * 12 + k];
            fp r = form ? fpd_pow_prog(a, C_PROG_PM3D4) : fp_pow_prog(a, C_PROG_PM3D4);
#pragma unroll
            for (int k = 0; k < 12; k++) out[(size_t)i * 12 + k] = r.l[k];
        }
        if (lane == 0) atomicAdd(&done_by_form[form], 32u);
    }
}

int main(int argc, char** argv) {
    const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : (1u << 19);
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    std::vector<uint32_t> h((size_t)n * 12);
    uint64_t x = 0x9e3779b97f4a7c15ull;
    for (size_t i = 0; i < h.size(); i++) {
        x ^= x << 13; x ^= x >> 7; x ^= x << 17;
        h[i] = (uint32_t)(x >> 11);
        if (i % 12 == 11) h[i] &= 0x0fffffffu;                  // < p (top limb of p is 0x1a0111ea)
    }
    uint32_t *d_in, *d_out[3];
    unsigned int* d_ctr;
    cudaMalloc(&d_in, h.size() * 4);
    for (int m = 0; m < 3; m++) cudaMalloc(&d_out[m], h.size() * 4);
    cudaMalloc(&d_ctr, 4 * (1 + 1024 + 2));
    cudaMemcpy(d_in, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float ms[3];
    unsigned int done[3][2];
    for (int rep = 0; rep < 2; rep++)
        for (int m = 0; m < 3; m++) {
            cudaMemset(d_ctr, 0, 4 * (1 + 1024 + 2));
            cudaEventRecord(e0);
            k_pow_mix<<<p.multiProcessorCount * 4, 128>>>(d_in, d_out[m], n, m, d_ctr, d_ctr + 1, d_ctr + 1025);
            cudaEventRecord(e1);
            cudaDeviceSynchronize();
            cudaEventElapsedTime(&ms[m], e0, e1);
            cudaMemcpy(done[m], d_ctr + 1025, 8, cudaMemcpyDeviceToHost);
        }
    cudaError_t err = cudaGetLastError();
    std::vector<uint32_t> o0(h.size()), o1(h.size()), o2(h.size());
    cudaMemcpy(o0.data(), d_out[0], h.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o1.data(), d_out[1], h.size() * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(o2.data(), d_out[2], h.size() * 4, cudaMemcpyDeviceToHost);
    size_t bad1 = 0, bad2 = 0;
    for (size_t i = 0; i < h.size(); i++) {
        bad1 += o0[i] != o1[i];
        bad2 += o0[i] != o2[i];
    }
    printf("{\"n\": %u, \"ms_int\": %.3f, \"ms_fp64\": %.3f, \"ms_mixed\": %.3f, \"mixed_done_int\": %u, \"mixed_done_fp64\": %u, "
           "\"mismatch_fp64\": %zu, \"mismatch_mixed\": %zu, \"cuda\": \"%s\"}\n",
           n, ms[0], ms[1], ms[2], done[2][0], done[2][1], bad1, bad2, cudaGetErrorString(err));
    return (bad1 || bad2) ? 1 : 0;
}
