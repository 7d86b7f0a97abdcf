// b200pos.cu -- C ABI (include/b200pos.h) over the kernels in kernels.cuh.  Single translation
// unit -> libb200pos.so (nvcc -gencode arch=compute_100a,code=sm_100a).  No CPU fallback anywhere.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <nvtx3/nvToolsExt.h>      // header-only NVTX v3: every C-ABI call is a named range in Nsight Systems / ncu --nvtx timelines

#include <algorithm>
#include <new>
#include <vector>

#include "../../include/b200pos.h"
#include "kernels.cuh"
#include "gather.cuh"
#include "rlc.cuh"

using namespace b2;

// dynamic shared memory of the tree phase in its shared-memory form: votes/prefix/weights u64[n+1], packed word u32[n], marks u32[n+1],
// and -- when it still fits -- the work list u32[n] of the marking phase
static size_t tree_smem_bytes(ghost_tree_args& A) {
    const size_t base = ((size_t)A.n + 1) * 8 + (size_t)A.n * 8 + 8;
    const size_t with_list = base + (size_t)A.n * 4;
    A.hard_list = with_list <= 226 * 1024 ? 1 : 0;
    return A.hard_list ? with_list : base;
}

struct dbuf {
    void* p = nullptr;
    size_t cap = 0;
};

struct vslot {
    dbuf haff, hflag, pkjac, pkst, saff, sflag, f, sumjac;
    dbuf g2aff, g2st;                                      // the slot's own decompressed signatures (pipelined epochs, B2_SEGSUM_TAIL)
    uint64_t n_sig_tail = 0;                               // signatures of the epoch whose segment sums the tail still has to run
    bool segsum_pending = false;
    dbuf pkjac_r, rscal, sg_aff, sg_flag, f_g, gpass;      // RLC batch mode: [r]PK, r, per-group signature sums, group Miller values / verdicts
    cudaEvent_t ev_join0 = nullptr, ev_join1 = nullptr, ev_seg = nullptr, ev_tail_done = nullptr, ev_fork = nullptr, ev_in = nullptr;
    cudaStream_t s_tail = nullptr;      // the slot's own tail stream: tails of consecutive epochs overlap each other
    // the slot's own side streams (hash-to-G2; pubkey aggregation + first Miller loop).  With ONE pair shared by all slots the
    // latency-bound side chains of consecutive epochs serialise (6 + 10 ms per epoch) -- invisible behind a 32 ms decompression,
    // but the bound of the step once an epoch is sharded over 8 GPUs and its decompression takes 3.7 ms.
    cudaStream_t s_aux[2] = {nullptr, nullptr};
};

struct b2_ctx {
    int device = 0;
    int n_sm = 148;
    cudaStream_t s_main = nullptr;
    // pipelined epochs: the decompression kernels of consecutive epochs alternate between two streams, so the blocks of epoch k+1 fill
    // the SMs that epoch k's last, partial wave leaves idle (a rank's share of a sharded epoch is 1.7 waves at N = 8: 5.04 ms instead of
    // 3.65 ms per 131 072 signatures when the kernels run back to back on one stream)
    cudaStream_t s_dec[4] = {nullptr, nullptr, nullptr, nullptr};
    unsigned n_dec_streams = 2;                       // B2_DEC_STREAMS (2..4)
    // Measured (profiles/r2d_emu_*): one rank's share at N = 8 / 4: 5.45 vs 5.97 and 9.75 vs 10.07 ms per epoch; at N = 1 (13.8 waves) the
    // overlap costs 0.3 ms instead, so the alternation is used only for decompressions of fewer than 6 waves (B2_DEC_ALTERNATE=0/1 forces)
    int dec_alternate = -1;                           // -1 auto, 0 never, 1 always
    unsigned dec_turn = 0;
    cudaEvent_t ev_votes_done = nullptr, ev_lmd_done = nullptr;
    vslot slot[B2_EPOCH_SLOTS];
    char err[512] = {0};
    uint64_t launches = 0;
    // registry
    uint64_t n_val = 0;
    uint32_t* d_records = nullptr;
    uint8_t* d_valid = nullptr;
    uint64_t* d_eff = nullptr;
    uint8_t* d_flags = nullptr;
    // latest messages
    unsigned long long* d_lmd_key = nullptr;
    uint32_t* d_lmd_block = nullptr;
    uint8_t* d_equiv = nullptr;
    // fork-choice variants: votes older than fc_min_key (epoch << 32) expire; v1.3 get_weight skips slashed validators
    // threads per block of the thread-per-aggregate kernels when they run under the epoch pipeline (see launch_miller)
    unsigned tail_block = 128;
    bool epoch_team = false;   // b2_epoch_set_pairing_form
    unsigned k2_block = 32;    // threads per aggregate of k_g1_aggregate under the epoch pipeline (B2_K2_BLOCK: 32/64/128)
    // K2 through the TMA-staged gather kernel (gather.cuh) instead of the LDG form; B2_K2_TMA=0 selects the LDG form (A/B, profiles/)
    bool k2_tma = true;
    // sliding-window tables of the decompression's exponentiations: B2_POW_SMEM=1 puts them in shared memory (halves the kernel's DRAM
    // traffic, 1.40 -> 0.73 GB per launch, ncu profiles/r2_g2_decompress_raw.csv) -- but the 48 KB per block compete with the tail
    // kernels' shared memory under the pipeline and the step is 0.6 ms SLOWER (36.11 vs 35.48 ms, profiles/r2_ab_*.json), so the
    // default keeps them in local memory
    bool pow_smem = false;
    // pipelined epochs: per-segment signature sums on the slot's tail stream instead of the caller's (B2_SEGSUM_TAIL=0: round-1 form)
    bool segsum_tail = true;
    // FastAggregateVerify mode: 0 = one pairing check per aggregate; 1 = random-linear-combination batches of B2_RLC_GROUP
    // aggregates with per-aggregate fallback (b2_set_verify_mode); d_rlc_seed = the verifier's secret 32-byte seed
    int verify_mode = 0;
    uint8_t* d_rlc_seed = nullptr;
    int pairing_form = 0;      // 0: team kernels for the synchronous calls, thread-per-aggregate under the pipeline; 1: always team; 2: always thread
    // threads per block of k_g2_decompress: its blocks fill the register file, so a smaller block is what a pairing warp of the
    // previous epoch displaces when the two overlap
    unsigned dec_block = 128;
    // SMs the persistent decompression kernel of the pipelined epoch path leaves empty for the pairing tails (kernels.cuh).
    // 0 = off (the default: measured slower); B2_RESERVE_SMS=n turns the experiment on.
    unsigned reserve_sms = 0;
    sm_mask reserved = {{0, 0, 0, 0}};
    unsigned long long* d_dec_counter = nullptr;      // one work counter per epoch slot
    // one-launch get_head (k_get_head_fused): ticket counter, and the result slot in mapped pinned host memory (head, sequence)
    bool head_fused = true;                           // B2_HEAD_FUSED=0: the two-launch + memcpy form of round 1
    unsigned int* d_ticket = nullptr;
    volatile uint32_t* h_head = nullptr;
    uint32_t* d_head_host = nullptr;
    uint32_t head_seq = 0;
    // multi-GPU get_head over NVLink peer memory (b2_fc_exchange_*): this rank's exchange block = accumulators 2 x n_blocks u64 followed by
    // flag rows 2 x B2_MAX_PEERS u32; peers' blocks opened through CUDA IPC
    void* fc_block = nullptr;
    size_t fc_block_bytes = 0;
    uint32_t fc_blocks_cap = 0;                       // n_blocks the block was sized for
    void* fc_peer[B2_MAX_PEERS] = {nullptr};
    int fc_rank = -1, fc_world = 0;
    uint32_t fc_seq = 0;
    unsigned long long* d_dbg = nullptr;              // clock64() stamps of the last get_head's phases (b2_debug_head_clocks)
    uint32_t* d_guard = nullptr;                      // device-side input guard word (kernels.cuh GuardBits), read by b2_guard_flags
    unsigned long long fc_min_key = 0;
    int fc_exclude_slashed = 0;
    // epoch participation flags (0 = current, 1 = previous) and the per-(validator, flag) election table
    uint32_t* d_part[2] = {nullptr, nullptr};
    uint32_t* d_part_first = nullptr;
    // block tree
    uint32_t n_blocks = 0;
    uint32_t *d_pre = nullptr, *d_inv = nullptr, *d_size_keep = nullptr, *d_rank = nullptr, *d_next = nullptr, *d_gsize = nullptr;
    bool packed_ok = false;
    uint32_t* d_packed = nullptr;     // size | rank << 15 | keep << 31 in pre-order (trees of < 32 768 blocks), else nullptr
    unsigned long long *d_votes = nullptr, *d_prefix = nullptr, *d_weight = nullptr, *d_w2 = nullptr;
    uint32_t* d_head = nullptr;
    // scratch
    dbuf sc_pkjac, sc_pkst, sc_haff, sc_hflag, sc_g2aff, sc_g2st, sc_rec, sc_val;
    dbuf in_a, in_b, in_c, in_d, in_e, in_f, in_g, out_a, out_b, sc_shuf, sc_pivot;
};

struct nvtx_scope {
    explicit nvtx_scope(const char* name) { nvtxRangePushA(name); }
    ~nvtx_scope() { nvtxRangePop(); }
};
#define B2_NVTX nvtx_scope nvtx_scope_(__func__)

static int fail_cuda(b2_ctx* c, cudaError_t e, const char* what) {
    if (c) snprintf(c->err, sizeof(c->err), "%s: %s", what, cudaGetErrorString(e));
    return B2_ECUDA;
}
#define CK(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return fail_cuda(ctx, e_, #call);   \
    } while (0)
#define CKL(ctx_)                                                          \
    do {                                                                   \
        cudaError_t e_ = cudaGetLastError();                               \
        if (e_ != cudaSuccess) return fail_cuda(ctx_, e_, "kernel launch"); \
        (ctx_)->launches++;                                                \
    } while (0)
#define REQUIRE(cond, msg)                                         \
    do {                                                           \
        if (!(cond)) {                                             \
            if (ctx) snprintf(ctx->err, sizeof(ctx->err), "%s", msg); \
            return B2_EINVAL;                                      \
        }                                                          \
    } while (0)

static int ensure(b2_ctx* ctx, dbuf& b, size_t bytes) {
    if (bytes <= b.cap && b.p) return B2_OK;
    if (b.p) CK(cudaFree(b.p));
    b.p = nullptr;
    b.cap = 0;
    size_t want = std::max<size_t>(bytes, 256);
    CK(cudaMalloc(&b.p, want));
    b.cap = want;
    return B2_OK;
}
template <class T> static int dev_alloc(b2_ctx* ctx, T** p, size_t count) {
    if (*p) {
        CK(cudaFree(*p));
        *p = nullptr;
    }
    CK(cudaMalloc((void**)p, std::max<size_t>(count * sizeof(T), 256)));
    return B2_OK;
}
#define B2_TEAM_SMEM (B2_TEAMS_PER_WARP * sizeof(team_ws))
static inline unsigned blocks_for(uint64_t n, unsigned t) { return (unsigned)((n + t - 1) / t); }

extern "C" {

int b2_init(int device, b2_ctx** out) {
    B2_NVTX;
    if (!out) return B2_EINVAL;
    *out = nullptr;
    // 8 pipeline slots x (tail + 2 side streams) + main/copy/fork-choice streams: more streams than the default 8 hardware queues,
    // whose aliasing creates false dependencies between slots.  Only effective when this is the first CUDA call of the process
    // (bench.py and the package __init__ also set it before torch initialises CUDA).
    setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return B2_ENODEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B2_ENODEVICE;
    if (prop.major < 10) return B2_ENODEVICE;      // sm_100a cubin only
    b2_ctx* ctx = new (std::nothrow) b2_ctx();
    if (!ctx) return B2_ENOMEM;
    ctx->device = device;
    if (cudaSetDevice(device) != cudaSuccess) {
        delete ctx;
        return B2_ENODEVICE;
    }
    // side streams get the highest priority: their kernels are small and latency-bound and must be scheduled as soon
    // as SM resources free up while a grid-filling kernel (signature decompression) occupies the main stream
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    cudaError_t e = cudaStreamCreateWithPriority(&ctx->s_main, cudaStreamNonBlocking, prio_lo);
    for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaStreamCreateWithPriority(&ctx->s_dec[i], cudaStreamNonBlocking, prio_lo);
    if (const char* ev = getenv("B2_DEC_STREAMS")) {
        unsigned v = (unsigned)atoi(ev);
        if (v >= 2 && v <= 4) ctx->n_dec_streams = v;
    }
    if (const char* ev = getenv("B2_DEC_ALTERNATE")) ctx->dec_alternate = atoi(ev) != 0 ? 1 : 0;
    for (int i = 0; i < B2_EPOCH_SLOTS && e == cudaSuccess; i++) {
        e = cudaStreamCreateWithPriority(&ctx->slot[i].s_tail, cudaStreamNonBlocking, prio_hi);
        for (int k = 0; k < 2 && e == cudaSuccess; k++) e = cudaStreamCreateWithPriority(&ctx->slot[i].s_aux[k], cudaStreamNonBlocking, prio_hi);
    }
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_lmd_done, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_votes_done, cudaEventDisableTiming);
    for (int i = 0; i < B2_EPOCH_SLOTS && e == cudaSuccess; i++) {
        cudaEvent_t* evs[6] = {&ctx->slot[i].ev_join0, &ctx->slot[i].ev_join1, &ctx->slot[i].ev_seg, &ctx->slot[i].ev_tail_done, &ctx->slot[i].ev_fork,
                               &ctx->slot[i].ev_in};
        for (int k = 0; k < 6 && e == cudaSuccess; k++) e = cudaEventCreateWithFlags(evs[k], cudaEventDisableTiming);
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ghost_tree, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_ghost_votes_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_g2_decompress, cudaFuncAttributeMaxDynamicSharedMemorySize, 432 * 128);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_g1_gather_tma<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)sizeof(gather_ws));
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_g1_gather_tma<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * (int)sizeof(gather_ws));
    ctx->n_sm = prop.multiProcessorCount;
    if (const char* e = getenv("B2_TAIL_BLOCK")) {          // tuning knob (32/64/128); the default is the measured best
        unsigned v = (unsigned)atoi(e);
        if (v == 32 || v == 64 || v == 128) ctx->tail_block = v;
    }
    if (const char* e = getenv("B2_K2_BLOCK")) {
        unsigned v = (unsigned)atoi(e);
        if (v == 32 || v == 64 || v == 128) ctx->k2_block = v;
    }
    if (const char* e = getenv("B2_K2_TMA")) ctx->k2_tma = atoi(e) != 0;
    if (const char* e = getenv("B2_POW_SMEM")) ctx->pow_smem = atoi(e) != 0;
    if (const char* e = getenv("B2_SEGSUM_TAIL")) ctx->segsum_tail = atoi(e) != 0;
    if (const char* e = getenv("B2_PAIRING_FORM")) ctx->pairing_form = !strcmp(e, "team") ? 1 : (!strcmp(e, "thread") ? 2 : 0);
    if (const char* e = getenv("B2_RESERVE_SMS")) {
        unsigned v = (unsigned)atoi(e);
        if (v < (unsigned)ctx->n_sm / 2) ctx->reserve_sms = v;
    }
    if (ctx->reserve_sms >= (unsigned)ctx->n_sm / 2) ctx->reserve_sms = ctx->n_sm / 9;
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_dec_counter, sizeof(unsigned long long) * B2_EPOCH_SLOTS);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_dbg, 32 * 8);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_dbg, 0, 32 * 8);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_ticket, 64);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_ticket, 0, 64);
    if (e == cudaSuccess) e = cudaHostAlloc((void**)&ctx->h_head, 64, cudaHostAllocMapped);
    if (e == cudaSuccess) {
        memset((void*)ctx->h_head, 0, 64);
        e = cudaHostGetDevicePointer((void**)&ctx->d_head_host, (void*)ctx->h_head, 0);
    }
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_get_head_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_get_head_fused_nvl, cudaFuncAttributeMaxDynamicSharedMemorySize, 226 * 1024);
    if (const char* ev = getenv("B2_HEAD_FUSED")) ctx->head_fused = atoi(ev) != 0;
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_rlc_seed, 32);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_rlc_seed, 0, 32);
    if (e == cudaSuccess) e = cudaMalloc(&ctx->d_guard, 256);
    if (e == cudaSuccess) e = cudaMemset(ctx->d_guard, 0, 256);
    if (e == cudaSuccess && ctx->reserve_sms > 0) {
        // the SM ids that exist (they need not be 0..n_sm-1); reserve the highest `reserve_sms` of them
        unsigned int* d_seen = nullptr;
        unsigned int seen[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        e = cudaMalloc(&d_seen, sizeof(seen));
        if (e == cudaSuccess) e = cudaMemset(d_seen, 0, sizeof(seen));
        if (e == cudaSuccess) {
            k_probe_smid<<<ctx->n_sm * 8, 64>>>(d_seen);
            e = cudaMemcpy(seen, d_seen, sizeof(seen), cudaMemcpyDeviceToHost);
        }
        if (d_seen) cudaFree(d_seen);
        unsigned left = ctx->reserve_sms, found = 0;
        for (int id = 255; id >= 0; id--) {
            if (!((seen[id >> 5] >> (id & 31)) & 1u)) continue;
            found++;
            if (left) {
                ctx->reserved.w[id >> 6] |= 1ull << (id & 63);
                left--;
            }
        }
        if (found < (unsigned)ctx->n_sm) {          // the probe did not reach every SM: do not reserve blindly
            ctx->reserved = sm_mask{{0, 0, 0, 0}};
            ctx->reserve_sms = 0;
        }
    }
    if (const char* e = getenv("B2_DEC_BLOCK")) {
        unsigned v = (unsigned)atoi(e);
        if (v == 32 || v == 64 || v == 128) ctx->dec_block = v;
    }
    if (e != cudaSuccess) {
        delete ctx;
        return B2_ECUDA;
    }
    *out = ctx;
    return B2_OK;
}

void b2_destroy(b2_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    void* ptrs[] = {ctx->d_packed, ctx->d_part[0], ctx->d_part[1], ctx->d_part_first, ctx->d_records, ctx->d_valid, ctx->d_eff, ctx->d_flags, ctx->d_lmd_key, ctx->d_lmd_block, ctx->d_equiv, ctx->d_pre,
                    ctx->d_inv, ctx->d_size_keep, ctx->d_gsize, ctx->d_w2, ctx->d_rank, ctx->d_next, ctx->d_votes, ctx->d_prefix,
                    ctx->d_weight, ctx->d_head};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    dbuf* bufs[] = {&ctx->sc_pkjac, &ctx->sc_pkst, &ctx->sc_haff, &ctx->sc_hflag,
                    &ctx->sc_g2aff, &ctx->sc_g2st, &ctx->sc_rec, &ctx->sc_val, &ctx->in_a, &ctx->in_b, &ctx->in_c, &ctx->in_d,
                    &ctx->in_e, &ctx->in_f, &ctx->in_g, &ctx->out_a, &ctx->out_b, &ctx->sc_shuf, &ctx->sc_pivot};
    for (dbuf* b : bufs)
        if (b->p) cudaFree(b->p);
    for (int i = 0; i < B2_EPOCH_SLOTS; i++) {
        vslot& V = ctx->slot[i];
        dbuf* vb[] = {&V.haff, &V.hflag, &V.pkjac, &V.pkst, &V.saff, &V.sflag, &V.f, &V.sumjac, &V.pkjac_r, &V.rscal, &V.sg_aff, &V.sg_flag, &V.f_g, &V.gpass, &V.g2aff, &V.g2st};
        for (dbuf* b : vb)
            if (b->p) cudaFree(b->p);
        cudaEvent_t evs[6] = {V.ev_join0, V.ev_join1, V.ev_seg, V.ev_tail_done, V.ev_fork, V.ev_in};
        for (cudaEvent_t ev : evs)
            if (ev) cudaEventDestroy(ev);
        if (V.s_tail) cudaStreamDestroy(V.s_tail);
        for (int k = 0; k < 2; k++)
            if (V.s_aux[k]) cudaStreamDestroy(V.s_aux[k]);
    }
    if (ctx->s_main) cudaStreamDestroy(ctx->s_main);
    for (int i = 0; i < 4; i++)
        if (ctx->s_dec[i]) cudaStreamDestroy(ctx->s_dec[i]);
    if (ctx->ev_lmd_done) cudaEventDestroy(ctx->ev_lmd_done);
    if (ctx->d_dec_counter) cudaFree(ctx->d_dec_counter);
    if (ctx->d_guard) cudaFree(ctx->d_guard);
    if (ctx->d_rlc_seed) cudaFree(ctx->d_rlc_seed);
    if (ctx->d_ticket) cudaFree(ctx->d_ticket);
    if (ctx->d_dbg) cudaFree(ctx->d_dbg);
    for (int r = 0; r < B2_MAX_PEERS; r++)
        if (ctx->fc_peer[r] && r != ctx->fc_rank) cudaIpcCloseMemHandle(ctx->fc_peer[r]);
    if (ctx->fc_block) cudaFree(ctx->fc_block);
    if (ctx->h_head) cudaFreeHost((void*)ctx->h_head);
    if (ctx->ev_votes_done) cudaEventDestroy(ctx->ev_votes_done);
    delete ctx;
}

const char* b2_last_error(b2_ctx* ctx) { return ctx ? ctx->err : "null context"; }
uint64_t b2_launch_count(b2_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint32_t b2_tree_size(b2_ctx* ctx) { return ctx ? ctx->n_blocks : 0; }
int b2_sync(b2_ctx* ctx) {
    B2_NVTX;
    REQUIRE(ctx, "null context");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    return B2_OK;
}

int b2_guard_flags(b2_ctx* ctx, uint32_t* flags_out) {
    B2_NVTX;
    REQUIRE(ctx && flags_out, "guard_flags: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    uint32_t f = 0;
    CK(cudaMemcpy(&f, ctx->d_guard, 4, cudaMemcpyDeviceToHost));
    CK(cudaMemset(ctx->d_guard, 0, 4));
    *flags_out = f;
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ registry
int b2_registry_load(b2_ctx* ctx, const uint8_t* pk48, const uint64_t* eff, const uint8_t* flags, uint64_t n, uint8_t* pk_valid_out) {
    B2_NVTX;
    REQUIRE(ctx && pk48 && eff && flags && n > 0 && n < (1ull << 32), "registry_load: bad arguments");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = dev_alloc(ctx, &ctx->d_records, n * 24)) || (rc = dev_alloc(ctx, &ctx->d_valid, n)) || (rc = dev_alloc(ctx, &ctx->d_eff, n)) ||
        (rc = dev_alloc(ctx, &ctx->d_flags, n)) || (rc = dev_alloc(ctx, &ctx->d_lmd_key, n)) || (rc = dev_alloc(ctx, &ctx->d_lmd_block, n)) ||
        (rc = dev_alloc(ctx, &ctx->d_equiv, n)))
        return rc;
    if ((rc = ensure(ctx, ctx->in_a, n * 48))) return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpyAsync(ctx->in_a.p, pk48, n * 48, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_eff, eff, n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_flags, flags, n, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_lmd_key, 0, n * 8, s));
    CK(cudaMemsetAsync(ctx->d_lmd_block, 0, n * 4, s));
    CK(cudaMemsetAsync(ctx->d_equiv, 0, n, s));
    k_registry_load<<<blocks_for(n, 128), 128, 0, s>>>((const uint8_t*)ctx->in_a.p, n, ctx->d_records, ctx->d_valid);
    CKL(ctx);
    if (pk_valid_out) CK(cudaMemcpyAsync(pk_valid_out, ctx->d_valid, n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (int w = 0; w < 2; w++)
        if (ctx->d_part[w]) {
            cudaFree(ctx->d_part[w]);
            ctx->d_part[w] = nullptr;
        }
    if (ctx->d_part_first) {
        cudaFree(ctx->d_part_first);
        ctx->d_part_first = nullptr;
    }
    ctx->n_val = n;
    return B2_OK;
}

// bls.KeyValidate for n explicit pubkeys, without touching the registry (scratch buffers only)
int b2_key_validate(b2_ctx* ctx, const uint8_t* pk48, uint64_t n, uint8_t* valid_out) {
    B2_NVTX;
    REQUIRE(ctx && pk48 && valid_out && n > 0 && n < (1ull << 32), "key_validate: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_a, n * 48)) || (rc = ensure(ctx, ctx->sc_rec, n * 96 + 16)) || (rc = ensure(ctx, ctx->sc_val, n + 16))) return rc;
    CK(cudaMemcpyAsync(ctx->in_a.p, pk48, n * 48, cudaMemcpyHostToDevice, s));
    k_g1_decompress_validate<<<blocks_for(n, 128), 128, 0, s>>>((const uint8_t*)ctx->in_a.p, n, (uint32_t*)ctx->sc_rec.p, (uint8_t*)ctx->sc_val.p);
    CKL(ctx);
    std::vector<uint8_t> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), ctx->sc_val.p, n, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(valid_out, tmp.data(), n);
    return B2_OK;
}

int b2_registry_update_balances(b2_ctx* ctx, const uint64_t* eff, const uint8_t* flags, uint64_t n) {
    B2_NVTX;
    REQUIRE(ctx && eff && flags && n == ctx->n_val && n > 0, "registry_update_balances: registry not loaded or size mismatch");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(ctx->d_eff, eff, n * 8, cudaMemcpyHostToDevice, ctx->s_main));
    CK(cudaMemcpyAsync(ctx->d_flags, flags, n, cudaMemcpyHostToDevice, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return B2_OK;
}

// host-side validation of a committee batch (indices must address the registry)
static int check_batch(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, uint32_t bits_stride, uint32_t n_agg, uint64_t limit) {
    REQUIRE(off[0] == 0, "off[0] must be 0");
    for (uint32_t a = 0; a < n_agg; a++) {
        REQUIRE(off[a + 1] >= off[a], "off must be non-decreasing");
        REQUIRE((uint64_t)(off[a + 1] - off[a]) <= (uint64_t)bits_stride * 8, "committee larger than bits_stride*8");
    }
    const uint32_t total = off[n_agg];
    for (uint32_t i = 0; i < total; i++) REQUIRE(members[i] < limit, "member index outside the registry");
    return B2_OK;
}

// upload members/off/bits into in_a/in_b/in_c
static int upload_batch(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride, uint32_t n_agg,
                        cudaStream_t s) {
    int rc;
    const uint32_t total = off[n_agg];
    if ((rc = ensure(ctx, ctx->in_a, (size_t)total * 4 + 4)) || (rc = ensure(ctx, ctx->in_b, (size_t)(n_agg + 1) * 4)) ||
        (rc = ensure(ctx, ctx->in_c, (size_t)n_agg * bits_stride + 16)))
        return rc;
    if (total) CK(cudaMemcpyAsync(ctx->in_a.p, members, (size_t)total * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_b.p, off, (size_t)(n_agg + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_c.p, bits, (size_t)n_agg * bits_stride, cudaMemcpyHostToDevice, s));
    return B2_OK;
}

// K2 launch: one warp per aggregate through the TMA-staged gather (wpb warps = aggregates per block), or the LDG form with
// `ldg_block` threads per aggregate
static int launch_k2(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride, uint32_t n_agg,
                     uint32_t* d_pkjac, uint8_t* d_pkst, unsigned wpb, unsigned ldg_block, cudaStream_t s) {
    if (ctx->k2_tma) {
        k_g1_gather_tma<false><<<blocks_for(n_agg, wpb), wpb * 32, wpb * sizeof(gather_ws), s>>>(ctx->d_records, d_members, d_off, d_bits, bits_stride, n_agg,
                                                                                             d_pkjac, d_pkst, ctx->n_val, ctx->d_guard, nullptr);
    } else {
        k_g1_aggregate<<<n_agg, ldg_block, 0, s>>>(ctx->d_records, ctx->d_valid, d_members, d_off, d_bits, bits_stride, n_agg, d_pkjac, d_pkst, ctx->n_val,
                                                   ctx->d_guard);
    }
    CKL(ctx);
    return B2_OK;
}

static int g1_aggregate_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                            uint32_t n_agg, cudaStream_t s) {
    int rc;
    if ((rc = ensure(ctx, ctx->sc_pkjac, (size_t)n_agg * 144)) || (rc = ensure(ctx, ctx->sc_pkst, n_agg))) return rc;
    return launch_k2(ctx, d_members, d_off, d_bits, bits_stride, n_agg, (uint32_t*)ctx->sc_pkjac.p, (uint8_t*)ctx->sc_pkst.p, 4, 128, s);
}

int b2_g1_aggregate(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride, uint32_t n_agg,
                    uint8_t* out48, uint8_t* status) {
    B2_NVTX;
    REQUIRE(ctx && members && off && bits && out48 && status && n_agg > 0 && bits_stride > 0, "g1_aggregate: bad arguments");
    REQUIRE(ctx->n_val > 0, "g1_aggregate: registry not loaded");
    int rc;
    if ((rc = check_batch(ctx, members, off, bits_stride, n_agg, ctx->n_val))) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    if ((rc = upload_batch(ctx, members, off, bits, bits_stride, n_agg, s))) return rc;
    if ((rc = g1_aggregate_dev(ctx, (const uint32_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (const uint8_t*)ctx->in_c.p, bits_stride, n_agg, s)))
        return rc;
    if ((rc = ensure(ctx, ctx->out_a, (size_t)n_agg * 48))) return rc;
    k_g1_compress<<<blocks_for(n_agg, 64), 64, 0, s>>>((const uint32_t*)ctx->sc_pkjac.p, n_agg, (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    std::vector<uint8_t> t48((size_t)n_agg * 48), tst(n_agg);
    CK(cudaMemcpyAsync(t48.data(), ctx->out_a.p, t48.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(tst.data(), ctx->sc_pkst.p, n_agg, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(out48, t48.data(), t48.size());
    memcpy(status, tst.data(), n_agg);
    return B2_OK;
}

// The gather stage of K2 ALONE (no additions): XOR checksum of the records each aggregate selects.  form 1: TMA-staged
// (k_g1_gather_tma<true>), form 0: plain 128-bit loads (k_g1_gather_ldg_probe).  bench.py times it for `roofline.gather`.
int b2_gather_probe_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride, uint32_t n_agg,
                        int form, uint32_t* d_checksum_out, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_members && d_off && d_bits && d_checksum_out && n_agg > 0 && bits_stride > 0 && (form == 0 || form == 1), "gather_probe_dev: bad arguments");
    REQUIRE(ctx->n_val > 0, "gather_probe: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (form == 1) {
        k_g1_gather_tma<true><<<blocks_for(n_agg, 4), 128, 4 * sizeof(gather_ws), s>>>(ctx->d_records, d_members, d_off, d_bits, bits_stride, n_agg, nullptr,
                                                                                    nullptr, ctx->n_val, ctx->d_guard, d_checksum_out);
    } else {
        k_g1_gather_ldg_probe<<<n_agg, 128, 0, s>>>(ctx->d_records, d_members, d_off, d_bits, bits_stride, n_agg, ctx->n_val, d_checksum_out);
    }
    CKL(ctx);
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ bls.Aggregate
// front: stage 1 + 2 (decompress every signature, per-segment Jacobian sums into V.sumjac);  finish: stage 3 (one
// inversion per segment -> compressed bytes; with `handoff` also the affine point + subgroup-checked flag for verify_main).
// `seg_stream`: the stream the per-segment sums run on.  Same as `s` for the synchronous entry points.  Under the epoch pipeline it
// is the slot's TAIL stream: the decompressed points then live in the slot's own buffer and the caller's stream goes straight on to
// the next epoch's decompression instead of waiting ~2 ms for a 2 048-warp kernel that cannot fill the chip.
static int aggregate_front(b2_ctx* ctx, vslot& V, const uint8_t* d_sig96, const uint32_t* d_seg_off, uint32_t n_seg, uint64_t n_sig,
                           int32_t* d_seg_status, cudaStream_t s, int reserve_slot = -1, bool own_buffers = false) {
    int rc;
    dbuf& aff = own_buffers ? V.g2aff : ctx->sc_g2aff;
    dbuf& st = own_buffers ? V.g2st : ctx->sc_g2st;
    if ((rc = ensure(ctx, aff, (size_t)n_sig * 192 + 16)) || (rc = ensure(ctx, st, n_sig + 16)) || (rc = ensure(ctx, V.sumjac, (size_t)n_seg * 288))) return rc;
    if (n_sig && reserve_slot >= 0 && ctx->reserve_sms > 0 && n_sig >= (uint64_t)ctx->n_sm * 512) {
        unsigned long long* ctr = ctx->d_dec_counter + reserve_slot;
        CK(cudaMemsetAsync(ctr, 0, sizeof(unsigned long long), s));
        k_g2_decompress_persistent<<<ctx->n_sm * 4, 128, 0, s>>>(d_sig96, n_sig, (uint32_t*)aff.p, (uint8_t*)st.p, ctr, ctx->reserved);
        CKL(ctx);
    } else if (n_sig) {
        k_g2_decompress<<<blocks_for(n_sig, ctx->dec_block), ctx->dec_block, ctx->pow_smem ? 432 * ctx->dec_block : 0, s>>>(
            d_sig96, n_sig, (uint32_t*)aff.p, (uint8_t*)st.p, ctx->pow_smem ? 1 : 0);
        CKL(ctx);
    }
    if (own_buffers) return B2_OK;          // the segment sums follow on the tail stream (aggregate_segments)
    k_g2_segment_sum<<<n_seg, 32, 0, s>>>((const uint32_t*)aff.p, (const uint8_t*)st.p, d_seg_off, n_seg, (uint32_t*)V.sumjac.p, d_seg_status, n_sig, ctx->d_guard);
    CKL(ctx);
    return B2_OK;
}
static int aggregate_segments(b2_ctx* ctx, vslot& V, const uint32_t* d_seg_off, uint32_t n_seg, uint64_t n_sig, int32_t* d_seg_status, cudaStream_t t) {
    k_g2_segment_sum<<<n_seg, 32, 0, t>>>((const uint32_t*)V.g2aff.p, (const uint8_t*)V.g2st.p, d_seg_off, n_seg, (uint32_t*)V.sumjac.p, d_seg_status, n_sig,
                                           ctx->d_guard);
    CKL(ctx);
    return B2_OK;
}
static int aggregate_finish(b2_ctx* ctx, vslot& V, uint32_t n_seg, uint8_t* d_out96, const int32_t* d_seg_status, bool handoff, cudaStream_t s,
                            unsigned tb = 32) {
    int rc;
    if (handoff && ((rc = ensure(ctx, V.saff, (size_t)n_seg * 192)) || (rc = ensure(ctx, V.sflag, n_seg)))) return rc;
    k_g2_finish<<<blocks_for(n_seg, tb), tb, 0, s>>>((const uint32_t*)V.sumjac.p, d_seg_status, n_seg, d_out96,
                                                    handoff ? (uint32_t*)V.saff.p : nullptr, handoff ? (uint8_t*)V.sflag.p : nullptr);
    CKL(ctx);
    return B2_OK;
}

int b2_aggregate_dev(b2_ctx* ctx, const uint8_t* d_sig96, const uint32_t* d_seg_off, uint32_t n_seg, uint64_t n_sig, uint8_t* d_out96,
                     int32_t* d_seg_status, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_seg_off && d_out96 && d_seg_status && n_seg > 0 && (n_sig == 0 || d_sig96), "aggregate_dev: bad arguments");
    CK(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = aggregate_front(ctx, ctx->slot[0], d_sig96, d_seg_off, n_seg, n_sig, d_seg_status, (cudaStream_t)stream))) return rc;
    return aggregate_finish(ctx, ctx->slot[0], n_seg, d_out96, d_seg_status, false, (cudaStream_t)stream);
}

int b2_aggregate(b2_ctx* ctx, const uint8_t* sig96, const uint32_t* seg_off, uint32_t n_seg, uint8_t* out96, int32_t* seg_status) {
    B2_NVTX;
    REQUIRE(ctx && seg_off && out96 && seg_status && n_seg > 0, "aggregate: bad arguments");
    REQUIRE(seg_off[0] == 0, "aggregate: seg_off[0] must be 0");
    for (uint32_t s = 0; s < n_seg; s++) REQUIRE(seg_off[s + 1] >= seg_off[s], "aggregate: seg_off must be non-decreasing");
    const uint64_t n_sig = seg_off[n_seg];
    REQUIRE(n_sig == 0 || sig96, "aggregate: null signatures");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_d, n_sig * 96 + 16)) || (rc = ensure(ctx, ctx->in_b, (size_t)(n_seg + 1) * 4)) ||
        (rc = ensure(ctx, ctx->out_a, (size_t)n_seg * 96)) || (rc = ensure(ctx, ctx->out_b, (size_t)n_seg * 4)))
        return rc;
    if (n_sig) CK(cudaMemcpyAsync(ctx->in_d.p, sig96, n_sig * 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_b.p, seg_off, (size_t)(n_seg + 1) * 4, cudaMemcpyHostToDevice, s));
    if ((rc = b2_aggregate_dev(ctx, (const uint8_t*)ctx->in_d.p, (const uint32_t*)ctx->in_b.p, n_seg, n_sig, (uint8_t*)ctx->out_a.p,
                               (int32_t*)ctx->out_b.p, s)))
        return rc;
    std::vector<uint8_t> t96((size_t)n_seg * 96);
    std::vector<int32_t> tst(n_seg);
    CK(cudaMemcpyAsync(t96.data(), ctx->out_a.p, t96.size(), cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(tst.data(), ctx->out_b.p, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(out96, t96.data(), t96.size());
    memcpy(seg_status, tst.data(), (size_t)n_seg * 4);
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ FastAggregateVerify
// Verification pipeline on three streams.  Nothing on the two side streams depends on the signatures:
//   aux0: K4 hash-to-G2
//   aux1: K2 pubkey gather+aggregate (or the explicit-key variant), then [after aux0] the Miller loop e(PK_agg, H(m))
//   main: signature decompression + subgroup check, Miller loop e(-g1, sig), [join aux1] final exponentiation.
// verify_fork() enqueues the side work forked from the *current* position of the caller's stream, verify_main()
// the signature half.  All intermediate buffers and events live in a `vslot`; the pipelined epoch API alternates
// between two slots so that the latency-bound tail of epoch k overlaps with the signature decompression of epoch k+1.
struct pk_source {
    const uint32_t *d_members, *d_off;
    const uint8_t* d_bits;
    uint32_t bits_stride;
    const uint8_t* d_pk48;      // explicit-key form when non-null (d_off = pk offsets)
    uint64_t n_pk;
};
// Two forms of K5/K6.  `team`: three lanes per pairing -- shortest critical path, used when the caller waits for the result
// (synchronous entry points).  Thread-per-item: 2.7x longer, but ~40 % fewer warp instructions in total -- used by the pipelined
// epoch API, where the latency hides behind the next epoch's decompression and only the stolen multiply-pipe time counts.
static inline bool use_team(const b2_ctx* ctx, bool team) { return ctx->pairing_form == 0 ? team : ctx->pairing_form == 1; }
static int launch_miller(b2_ctx* ctx, vslot& V, uint32_t n_agg, int mode, bool team, cudaStream_t s) {
    team = use_team(ctx, team);
    if (team) {
        k_miller_team<<<blocks_for(n_agg, B2_TEAMS_PER_WARP), 32, B2_TEAM_SMEM, s>>>(
            (const uint32_t*)V.pkjac.p, (const uint8_t*)V.pkst.p, (const uint32_t*)V.haff.p, (const uint8_t*)V.hflag.p, (const uint32_t*)V.saff.p,
            (const uint8_t*)V.sflag.p, n_agg, (uint32_t*)V.f.p, mode);
    } else {
        k_miller<<<blocks_for(n_agg, ctx->tail_block), ctx->tail_block, 0, s>>>((const uint32_t*)V.pkjac.p, (const uint8_t*)V.pkst.p, (const uint32_t*)V.haff.p,
                                                      (const uint8_t*)V.hflag.p, (const uint32_t*)V.saff.p, (const uint8_t*)V.sflag.p, n_agg,
                                                      (uint32_t*)V.f.p, mode);
    }
    CKL(ctx);
    return B2_OK;
}
static int launch_final(b2_ctx* ctx, vslot& V, uint32_t n_agg, uint8_t* d_ok, bool team, cudaStream_t s) {
    team = use_team(ctx, team);
    if (team) {
        k_final_team<<<blocks_for(n_agg, B2_TEAMS_PER_WARP), 32, B2_TEAM_SMEM, s>>>((const uint32_t*)V.f.p, (const uint8_t*)V.pkst.p,
                                                                                  (const uint8_t*)V.sflag.p, n_agg, d_ok);
    } else {
        k_final_verdict<<<blocks_for(n_agg, ctx->tail_block), ctx->tail_block, 0, s>>>((const uint32_t*)V.f.p, (const uint8_t*)V.pkst.p, (const uint8_t*)V.sflag.p, n_agg, d_ok);
    }
    CKL(ctx);
    return B2_OK;
}
static int verify_fork(b2_ctx* ctx, vslot& V, const pk_source& P, const uint8_t* d_msg32, uint32_t n_agg, cudaStream_t s, bool team = true) {
    int rc;
    if ((rc = ensure(ctx, V.haff, (size_t)n_agg * 192)) || (rc = ensure(ctx, V.hflag, n_agg)) || (rc = ensure(ctx, V.saff, (size_t)n_agg * 192)) ||
        (rc = ensure(ctx, V.sflag, n_agg)) || (rc = ensure(ctx, V.f, (size_t)n_agg * 2 * 576)) || (rc = ensure(ctx, V.pkjac, (size_t)n_agg * 144)) ||
        (rc = ensure(ctx, V.pkst, n_agg)))
        return rc;
    if (P.d_pk48 && ((rc = ensure(ctx, ctx->sc_rec, P.n_pk * 96 + 16)) || (rc = ensure(ctx, ctx->sc_val, P.n_pk + 16)))) return rc;
    CK(cudaEventRecord(V.ev_fork, s));
    CK(cudaStreamWaitEvent(V.s_aux[0], V.ev_fork, 0));
    CK(cudaStreamWaitEvent(V.s_aux[1], V.ev_fork, 0));
    const unsigned tb = team ? 32u : ctx->tail_block;
    k_hash_to_g2<<<blocks_for(2ull * n_agg, tb), tb, 0, V.s_aux[0]>>>(d_msg32, n_agg, (uint32_t*)V.haff.p, (uint8_t*)V.hflag.p);
    CKL(ctx);
    CK(cudaEventRecord(V.ev_join0, V.s_aux[0]));
    if (P.d_pk48) {
        if (P.n_pk) {
            k_g1_decompress_validate<<<blocks_for(P.n_pk, 128), 128, 0, V.s_aux[1]>>>(P.d_pk48, P.n_pk, (uint32_t*)ctx->sc_rec.p, (uint8_t*)ctx->sc_val.p);
            CKL(ctx);
        }
        k_g1_segment_sum<<<n_agg, 128, 0, V.s_aux[1]>>>((const uint32_t*)ctx->sc_rec.p, (const uint8_t*)ctx->sc_val.p, P.d_off, n_agg,
                                                          (uint32_t*)V.pkjac.p, (uint8_t*)V.pkst.p);
        CKL(ctx);
    } else {
        // one block per aggregate.  128 threads when the caller waits; ONE WARP under the epoch pipeline: the 5-round shuffle tree costs
        // every warp 80 Fp multiplications whatever its share of the 512 records, so 4 warps spend 2.1x the multiply-pipe time of one
        const unsigned k2_block = team ? 128u : ctx->k2_block;
        if ((rc = launch_k2(ctx, P.d_members, P.d_off, P.d_bits, P.bits_stride, n_agg, (uint32_t*)V.pkjac.p, (uint8_t*)V.pkst.p, team ? 4u : 1u, k2_block,
                            V.s_aux[1])))
            return rc;
    }
    if (ctx->verify_mode == 1) {
        // RLC: the pubkey-side Miller loop runs on [r_i] PK_i (thread-per-aggregate kernels: this mode trades latency for work)
        const uint32_t n_groups = blocks_for(n_agg, B2_RLC_GROUP);
        if ((rc = ensure(ctx, V.pkjac_r, (size_t)n_agg * 144)) || (rc = ensure(ctx, V.rscal, (size_t)n_agg * 8)) || (rc = ensure(ctx, V.sg_aff, (size_t)n_groups * 192)) ||
            (rc = ensure(ctx, V.sg_flag, n_groups)) || (rc = ensure(ctx, V.f_g, (size_t)n_groups * 2 * 576)) || (rc = ensure(ctx, V.gpass, n_groups)))
            return rc;
        k_rlc_pk<<<blocks_for(n_agg, ctx->tail_block), ctx->tail_block, 0, V.s_aux[1]>>>(ctx->d_rlc_seed, d_msg32, (const uint32_t*)V.pkjac.p, (const uint8_t*)V.pkst.p,
                                                                                        n_agg, (unsigned long long*)V.rscal.p, (uint32_t*)V.pkjac_r.p);
        CKL(ctx);
        CK(cudaStreamWaitEvent(V.s_aux[1], V.ev_join0, 0));
        k_miller<<<blocks_for(n_agg, ctx->tail_block), ctx->tail_block, 0, V.s_aux[1]>>>((const uint32_t*)V.pkjac_r.p, (const uint8_t*)V.pkst.p, (const uint32_t*)V.haff.p,
                                                                                        (const uint8_t*)V.hflag.p, nullptr, nullptr, n_agg, (uint32_t*)V.f.p, 1);
        CKL(ctx);
        CK(cudaEventRecord(V.ev_join1, V.s_aux[1]));
        return B2_OK;
    }
    CK(cudaStreamWaitEvent(V.s_aux[1], V.ev_join0, 0));
    if ((rc = launch_miller(ctx, V, n_agg, 1, team, V.s_aux[1]))) return rc;
    CK(cudaEventRecord(V.ev_join1, V.s_aux[1]));
    return B2_OK;
}
// RLC batch mode, signature half + verdicts (V.saff / V.sflag hold the affine, subgroup-checked aggregate signatures)
static int verify_main_rlc(b2_ctx* ctx, vslot& V, uint32_t n_agg, uint8_t* d_ok, cudaStream_t s) {
    const uint32_t n_groups = blocks_for(n_agg, B2_RLC_GROUP);
    const unsigned tb = ctx->tail_block;
    const uint8_t *pkst = (const uint8_t*)V.pkst.p, *sflag = (const uint8_t*)V.sflag.p;
    k_rlc_sig<<<n_groups, 32, 0, s>>>((const uint32_t*)V.saff.p, sflag, pkst, (const unsigned long long*)V.rscal.p, n_agg, (uint32_t*)V.sg_aff.p, (uint8_t*)V.sg_flag.p);
    CKL(ctx);
    // e(-g1, sum [r_i] S_i) per group
    k_miller<<<blocks_for(n_groups, tb), tb, 0, s>>>(nullptr, nullptr, nullptr, nullptr, (const uint32_t*)V.sg_aff.p, (const uint8_t*)V.sg_flag.p, n_groups, (uint32_t*)V.f_g.p, 2);
    CKL(ctx);
    CK(cudaStreamWaitEvent(s, V.ev_join1, 0));
    k_rlc_final<<<blocks_for(n_groups, 32), 32, 0, s>>>((const uint32_t*)V.f.p, (const uint32_t*)V.f_g.p, pkst, sflag, n_agg, n_groups, (uint8_t*)V.gpass.p);
    CKL(ctx);
    k_rlc_verdict<<<blocks_for(n_agg, 128), 128, 0, s>>>(pkst, sflag, (const uint8_t*)V.gpass.p, n_agg, d_ok);
    CKL(ctx);
    // fallback: the members of a failed group, aggregate by aggregate (threads of passed groups exit at once)
    k_miller<<<blocks_for(2 * n_agg, tb), tb, 0, s>>>((const uint32_t*)V.pkjac.p, pkst, (const uint32_t*)V.haff.p, (const uint8_t*)V.hflag.p, (const uint32_t*)V.saff.p, sflag,
                                                       n_agg, (uint32_t*)V.f.p, 0, (const uint8_t*)V.gpass.p, B2_RLC_GROUP);
    CKL(ctx);
    k_final_verdict<<<blocks_for(n_agg, tb), tb, 0, s>>>((const uint32_t*)V.f.p, pkst, sflag, n_agg, d_ok, (const uint8_t*)V.gpass.p, B2_RLC_GROUP);
    CKL(ctx);
    return B2_OK;
}
// d_sig96 == nullptr: the signature points are already in V.saff / V.sflag (handed over by aggregate_finish)
static int verify_main(b2_ctx* ctx, vslot& V, const uint8_t* d_sig96, uint32_t n_agg, uint8_t* d_ok, cudaStream_t s, bool team = true) {
    int rc;
    if (d_sig96) {
        k_sig_prepare<<<blocks_for(n_agg, 32), 32, 0, s>>>(d_sig96, n_agg, (uint32_t*)V.saff.p, (uint8_t*)V.sflag.p);
        CKL(ctx);
    }
    if (ctx->verify_mode == 1) return verify_main_rlc(ctx, V, n_agg, d_ok, s);
    if ((rc = launch_miller(ctx, V, n_agg, 2, team, s))) return rc;
    CK(cudaStreamWaitEvent(s, V.ev_join1, 0));
    return launch_final(ctx, V, n_agg, d_ok, team, s);
}

int b2_fast_aggregate_verify_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                                 const uint8_t* d_msg32, const uint8_t* d_sig96, uint32_t n_agg, uint8_t* d_ok_out, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_members && d_off && d_bits && d_msg32 && d_sig96 && d_ok_out && n_agg > 0 && bits_stride > 0, "fast_aggregate_verify_dev: bad arguments");
    REQUIRE(ctx->n_val > 0, "fast_aggregate_verify: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    pk_source P = {d_members, d_off, d_bits, bits_stride, nullptr, 0};
    if ((rc = verify_fork(ctx, ctx->slot[0], P, d_msg32, n_agg, s))) return rc;
    return verify_main(ctx, ctx->slot[0], d_sig96, n_agg, d_ok_out, s);
}

// ---- one epoch for the validators of this rank: per-committee bls.Aggregate of the individual signatures,
// FastAggregateVerify of the aggregates, update_latest_messages for the accepted ones.
//   epoch_start: fork the signature-independent half (hash, pubkey aggregation, first Miller loop) on the side streams,
//                then decompress + segment-sum the signatures on the caller's stream (the grid-filling part);
//   epoch_tail : inversion/compress + subgroup check, second Miller loop, final exponentiation, LMD update -- latency-
//                bound work on few SMs; on `tail_stream` (the caller's own stream, or the context's tail stream so that
//                it overlaps with the next epoch's epoch_start in the other slot).
static int epoch_start(b2_ctx* ctx, int slot, const uint8_t* d_sig96, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                       uint32_t bits_stride, const uint8_t* d_msg32, uint32_t n_agg, uint64_t n_sig, int32_t* d_agg_status, cudaStream_t s, bool team,
                       bool own_buffers = false) {
    vslot& V = ctx->slot[slot];
    int rc;
    CK(cudaStreamWaitEvent(s, V.ev_tail_done, 0));      // the slot's previous user (B2_EPOCH_SLOTS-or-fewer epochs ago) must have drained
    pk_source P = {d_members, d_off, d_bits, bits_stride, nullptr, 0};
    if ((rc = verify_fork(ctx, V, P, d_msg32, n_agg, s, team))) return rc;
    V.segsum_pending = own_buffers;
    V.n_sig_tail = n_sig;
    cudaStream_t sd = s;
    const bool few_waves = n_sig < (uint64_t)6 * ctx->n_sm * 4 * ctx->dec_block;
    if (own_buffers && (ctx->dec_alternate == 1 || (ctx->dec_alternate < 0 && few_waves))) {      // the slot's own point buffer makes this legal
        sd = ctx->s_dec[ctx->dec_turn++ % ctx->n_dec_streams];
        CK(cudaEventRecord(V.ev_in, s));                // inputs ready, slot drained
        CK(cudaStreamWaitEvent(sd, V.ev_in, 0));
    }
    if ((rc = aggregate_front(ctx, V, d_sig96, d_off, n_agg, n_sig, d_agg_status, sd, team ? -1 : slot, own_buffers))) return rc;
    CK(cudaEventRecord(V.ev_seg, sd));
    return B2_OK;
}
static int epoch_tail(b2_ctx* ctx, int slot, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                      const uint64_t* d_target_epoch, const uint32_t* d_block_idx, uint32_t n_agg, uint8_t* d_agg_sig96, int32_t* d_agg_status,
                      uint8_t* d_ok_out, cudaStream_t t, bool team) {
    vslot& V = ctx->slot[slot];
    int rc;
    CK(cudaStreamWaitEvent(t, V.ev_seg, 0));
    if (V.segsum_pending) {
        if ((rc = aggregate_segments(ctx, V, d_off, n_agg, V.n_sig_tail, d_agg_status, t))) return rc;
        V.segsum_pending = false;
    }
    if ((rc = aggregate_finish(ctx, V, n_agg, d_agg_sig96, d_agg_status, true, t, team ? 32u : ctx->tail_block))) return rc;
    if ((rc = verify_main(ctx, V, nullptr, n_agg, d_ok_out, t, team))) return rc;
    if (d_target_epoch) {
        CK(cudaStreamWaitEvent(t, ctx->ev_votes_done, 0));  // do not move the LMD table under a vote scatter that is still reading it
        CK(cudaStreamWaitEvent(t, ctx->ev_lmd_done, 0));    // latest messages are applied in epoch order even when tails overlap
        if ((rc = b2_latest_messages_update_dev(ctx, d_members, d_off, d_bits, bits_stride, d_target_epoch, d_block_idx, d_ok_out, n_agg, t))) return rc;
        CK(cudaEventRecord(ctx->ev_lmd_done, t));
    }   // else: the caller applies update_latest_messages itself, after exchanging the verdicts of all ranks (sharded epoch)
    CK(cudaEventRecord(V.ev_tail_done, t));
    return B2_OK;
}

int b2_epoch_dev(b2_ctx* ctx, const uint8_t* d_sig96, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                 uint32_t bits_stride, const uint8_t* d_msg32, const uint64_t* d_target_epoch, const uint32_t* d_block_idx, uint32_t n_agg,
                 uint64_t n_sig, uint8_t* d_agg_sig96, int32_t* d_agg_status, uint8_t* d_ok_out, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_sig96 && d_members && d_off && d_bits && d_msg32 && (!d_target_epoch == !d_block_idx) && d_agg_sig96 && d_agg_status && d_ok_out &&
                n_agg > 0 && bits_stride > 0, "epoch_dev: bad arguments");
    REQUIRE(ctx->n_val > 0, "epoch_dev: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    int rc;
    if ((rc = epoch_start(ctx, 0, d_sig96, d_members, d_off, d_bits, bits_stride, d_msg32, n_agg, n_sig, d_agg_status, s, true))) return rc;
    return epoch_tail(ctx, 0, d_members, d_off, d_bits, bits_stride, d_target_epoch, d_block_idx, n_agg, d_agg_sig96, d_agg_status, d_ok_out, s, true);
}

// Pipelined form: per epoch k, slot = k mod depth (depth <= B2_EPOCH_SLOTS): start(slot) on the caller's stream, tail(slot) on the
// slot's own high-priority stream, then the fork choice of the epoch on any stream after b2_epoch_wait_dev(slot, stream).  Tails of
// consecutive epochs overlap each other and the following decompressions; latest messages are still applied in epoch order.
int b2_epoch_start_dev(b2_ctx* ctx, int slot, const uint8_t* d_sig96, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits,
                       uint32_t bits_stride, const uint8_t* d_msg32, uint32_t n_agg, uint64_t n_sig, int32_t* d_agg_status, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && slot >= 0 && slot < B2_EPOCH_SLOTS && d_sig96 && d_members && d_off && d_bits && d_msg32 && d_agg_status && n_agg > 0 && bits_stride > 0,
            "epoch_start_dev: bad arguments");
    REQUIRE(ctx->n_val > 0, "epoch_start_dev: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    return epoch_start(ctx, slot, d_sig96, d_members, d_off, d_bits, bits_stride, d_msg32, n_agg, n_sig, d_agg_status, (cudaStream_t)stream,
                       ctx->epoch_team, ctx->segsum_tail);
}
int b2_epoch_tail_dev(b2_ctx* ctx, int slot, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                      const uint64_t* d_target_epoch, const uint32_t* d_block_idx, uint32_t n_agg, uint8_t* d_agg_sig96, int32_t* d_agg_status,
                      uint8_t* d_ok_out) {
    B2_NVTX;
    REQUIRE(ctx && slot >= 0 && slot < B2_EPOCH_SLOTS && d_members && d_off && d_bits && (!d_target_epoch == !d_block_idx) && d_agg_sig96 && d_agg_status && d_ok_out &&
                n_agg > 0 && bits_stride > 0, "epoch_tail_dev: bad arguments");
    CK(cudaSetDevice(ctx->device));
    return epoch_tail(ctx, slot, d_members, d_off, d_bits, bits_stride, d_target_epoch, d_block_idx, n_agg, d_agg_sig96, d_agg_status, d_ok_out,
                      ctx->slot[slot].s_tail, ctx->epoch_team);
}
int b2_epoch_set_pairing_form(b2_ctx* ctx, int form) {
    B2_NVTX;
    REQUIRE(ctx && (form == 0 || form == 1), "epoch_set_pairing_form: bad arguments");
    ctx->epoch_team = form == 1;
    return B2_OK;
}
int b2_epoch_wait_dev(b2_ctx* ctx, int slot, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && slot >= 0 && slot < B2_EPOCH_SLOTS, "epoch_wait_dev: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamWaitEvent((cudaStream_t)stream, ctx->slot[slot].ev_tail_done, 0));
    return B2_OK;
}

int b2_fast_aggregate_verify(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                             const uint8_t* msg32, const uint8_t* sig96, uint32_t n_agg, uint8_t* ok_out) {
    B2_NVTX;
    REQUIRE(ctx && members && off && bits && msg32 && sig96 && ok_out && n_agg > 0 && bits_stride > 0, "fast_aggregate_verify: bad arguments");
    REQUIRE(ctx->n_val > 0, "fast_aggregate_verify: registry not loaded");
    int rc;
    if ((rc = check_batch(ctx, members, off, bits_stride, n_agg, ctx->n_val))) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    if ((rc = upload_batch(ctx, members, off, bits, bits_stride, n_agg, s))) return rc;
    if ((rc = ensure(ctx, ctx->in_d, (size_t)n_agg * 96)) || (rc = ensure(ctx, ctx->in_e, (size_t)n_agg * 32)) || (rc = ensure(ctx, ctx->out_a, n_agg)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_d.p, sig96, (size_t)n_agg * 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_e.p, msg32, (size_t)n_agg * 32, cudaMemcpyHostToDevice, s));
    if ((rc = b2_fast_aggregate_verify_dev(ctx, (const uint32_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (const uint8_t*)ctx->in_c.p, bits_stride,
                                           (const uint8_t*)ctx->in_e.p, (const uint8_t*)ctx->in_d.p, n_agg, (uint8_t*)ctx->out_a.p, s)))
        return rc;
    std::vector<uint8_t> tok(n_agg);
    CK(cudaMemcpyAsync(tok.data(), ctx->out_a.p, n_agg, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(ok_out, tok.data(), n_agg);
    return B2_OK;
}

int b2_fast_aggregate_verify_pks(b2_ctx* ctx, const uint8_t* pk48, const uint32_t* pk_off, const uint8_t* msg32, const uint8_t* sig96,
                                 uint32_t n_agg, uint8_t* ok_out) {
    B2_NVTX;
    REQUIRE(ctx && pk_off && msg32 && sig96 && ok_out && n_agg > 0, "fast_aggregate_verify_pks: bad arguments");
    REQUIRE(pk_off[0] == 0, "pk_off[0] must be 0");
    for (uint32_t a = 0; a < n_agg; a++) REQUIRE(pk_off[a + 1] >= pk_off[a], "pk_off must be non-decreasing");
    const uint64_t n_pk = pk_off[n_agg];
    REQUIRE(n_pk == 0 || pk48, "null pubkeys");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_a, n_pk * 48 + 16)) || (rc = ensure(ctx, ctx->in_b, (size_t)(n_agg + 1) * 4)) ||
        (rc = ensure(ctx, ctx->in_d, (size_t)n_agg * 96)) || (rc = ensure(ctx, ctx->in_e, (size_t)n_agg * 32)) || (rc = ensure(ctx, ctx->out_a, n_agg)) ||
        (rc = ensure(ctx, ctx->sc_rec, n_pk * 96 + 16)) || (rc = ensure(ctx, ctx->sc_val, n_pk + 16)) ||
        (rc = ensure(ctx, ctx->sc_pkjac, (size_t)n_agg * 144)) || (rc = ensure(ctx, ctx->sc_pkst, n_agg)))
        return rc;
    if (n_pk) CK(cudaMemcpyAsync(ctx->in_a.p, pk48, n_pk * 48, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_b.p, pk_off, (size_t)(n_agg + 1) * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_d.p, sig96, (size_t)n_agg * 96, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_e.p, msg32, (size_t)n_agg * 32, cudaMemcpyHostToDevice, s));
    pk_source P = {nullptr, (const uint32_t*)ctx->in_b.p, nullptr, 0, (const uint8_t*)ctx->in_a.p, n_pk};
    if ((rc = verify_fork(ctx, ctx->slot[0], P, (const uint8_t*)ctx->in_e.p, n_agg, s))) return rc;
    if ((rc = verify_main(ctx, ctx->slot[0], (const uint8_t*)ctx->in_d.p, n_agg, (uint8_t*)ctx->out_a.p, s))) return rc;
    std::vector<uint8_t> tok(n_agg);
    CK(cudaMemcpyAsync(tok.data(), ctx->out_a.p, n_agg, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(ok_out, tok.data(), n_agg);
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ SkToPk / Sign / hash_to_g2
int b2_sk_to_pk(b2_ctx* ctx, const uint32_t* sk8, uint64_t n, uint8_t* pk48_out) {
    B2_NVTX;
    REQUIRE(ctx && sk8 && pk48_out && n > 0, "sk_to_pk: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_f, n * 32)) || (rc = ensure(ctx, ctx->out_a, n * 48))) return rc;
    CK(cudaMemcpyAsync(ctx->in_f.p, sk8, n * 32, cudaMemcpyHostToDevice, s));
    k_sk_to_pk<<<blocks_for(n, 64), 64, 0, s>>>((const uint32_t*)ctx->in_f.p, n, (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(pk48_out, ctx->out_a.p, n * 48, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

int b2_hash_to_g2(b2_ctx* ctx, const uint8_t* msg32, uint32_t n_msg, uint8_t* out96) {
    B2_NVTX;
    REQUIRE(ctx && msg32 && out96 && n_msg > 0, "hash_to_g2: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_e, (size_t)n_msg * 32)) || (rc = ensure(ctx, ctx->sc_haff, (size_t)n_msg * 192)) ||
        (rc = ensure(ctx, ctx->sc_hflag, n_msg)) || (rc = ensure(ctx, ctx->out_a, (size_t)n_msg * 96)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_e.p, msg32, (size_t)n_msg * 32, cudaMemcpyHostToDevice, s));
    k_hash_to_g2<<<blocks_for(2ull * n_msg, 32), 32, 0, s>>>((const uint8_t*)ctx->in_e.p, n_msg, (uint32_t*)ctx->sc_haff.p, (uint8_t*)ctx->sc_hflag.p);
    CKL(ctx);
    k_g2_compress_aff<<<blocks_for(n_msg, 64), 64, 0, s>>>((const uint32_t*)ctx->sc_haff.p, (const uint8_t*)ctx->sc_hflag.p, n_msg, (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(out96, ctx->out_a.p, (size_t)n_msg * 96, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

int b2_sign(b2_ctx* ctx, const uint32_t* sk8, const uint32_t* msg_idx, uint64_t n, const uint8_t* msg32, uint32_t n_msg, uint8_t* sig96_out) {
    B2_NVTX;
    REQUIRE(ctx && sk8 && msg_idx && msg32 && sig96_out && n > 0 && n_msg > 0, "sign: bad arguments");
    for (uint64_t i = 0; i < n; i++) REQUIRE(msg_idx[i] < n_msg, "sign: msg_idx out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_f, n * 32)) || (rc = ensure(ctx, ctx->in_g, n * 4)) || (rc = ensure(ctx, ctx->in_e, (size_t)n_msg * 32)) ||
        (rc = ensure(ctx, ctx->sc_haff, (size_t)n_msg * 192)) || (rc = ensure(ctx, ctx->sc_hflag, n_msg)) || (rc = ensure(ctx, ctx->out_a, n * 96)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_f.p, sk8, n * 32, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_g.p, msg_idx, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_e.p, msg32, (size_t)n_msg * 32, cudaMemcpyHostToDevice, s));
    k_hash_to_g2<<<blocks_for(2ull * n_msg, 32), 32, 0, s>>>((const uint8_t*)ctx->in_e.p, n_msg, (uint32_t*)ctx->sc_haff.p, (uint8_t*)ctx->sc_hflag.p);
    CKL(ctx);
    k_sign<<<blocks_for(n, 64), 64, 0, s>>>((const uint32_t*)ctx->in_f.p, (const uint32_t*)ctx->in_g.p, n, (const uint32_t*)ctx->sc_haff.p, (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(sig96_out, ctx->out_a.p, n * 96, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ batched SHA-256
int b2_sha256_batch(b2_ctx* ctx, const uint8_t* msgs, uint32_t msg_len, uint64_t n, uint8_t* out32) {
    B2_NVTX;
    REQUIRE(ctx && out32 && n > 0 && (msg_len == 0 || msgs), "sha256_batch: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_a, (size_t)n * msg_len + 16)) || (rc = ensure(ctx, ctx->out_a, (size_t)n * 32))) return rc;
    if (msg_len) CK(cudaMemcpyAsync(ctx->in_a.p, msgs, (size_t)n * msg_len, cudaMemcpyHostToDevice, s));
    k_sha256_fixed<<<blocks_for(n, 128), 128, 0, s>>>((const uint8_t*)ctx->in_a.p, msg_len, n, (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(out32, ctx->out_a.p, (size_t)n * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ SSZ signing roots
int b2_signing_roots(b2_ctx* ctx, const uint8_t* data128, const uint8_t* domain32, int per_attestation_domain, uint32_t n, uint8_t* out32) {
    B2_NVTX;
    REQUIRE(ctx && data128 && domain32 && out32 && n > 0, "signing_roots: bad arguments");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    const size_t dom_bytes = per_attestation_domain ? (size_t)n * 32 : 32;
    if ((rc = ensure(ctx, ctx->in_a, (size_t)n * 128)) || (rc = ensure(ctx, ctx->in_e, dom_bytes)) || (rc = ensure(ctx, ctx->out_a, (size_t)n * 32))) return rc;
    CK(cudaMemcpyAsync(ctx->in_a.p, data128, (size_t)n * 128, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_e.p, domain32, dom_bytes, cudaMemcpyHostToDevice, s));
    k_signing_roots<<<blocks_for(n, 64), 64, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint8_t*)ctx->in_e.p, per_attestation_domain ? 32u : 0u, n,
                                                    (uint8_t*)ctx->out_a.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(out32, ctx->out_a.p, (size_t)n * 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ SSZ wire decode (:714-717)
int b2_attestations_decode(b2_ctx* ctx, const uint8_t* wire, const uint32_t* woff, uint32_t n, uint32_t bits_stride, uint32_t max_bits,
                           uint8_t* bits_out, uint32_t* bit_len_out, uint8_t* data128_out, uint8_t* sig96_out, int32_t* status_out) {
    B2_NVTX;
    REQUIRE(ctx && wire && woff && n > 0 && bits_stride > 0 && bits_out && bit_len_out && data128_out && sig96_out && status_out,
            "attestations_decode: bad arguments");
    REQUIRE(woff[0] == 0, "attestations_decode: woff[0] must be 0");
    for (uint32_t i = 0; i < n; i++) REQUIRE(woff[i + 1] >= woff[i], "attestations_decode: woff must be non-decreasing");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    const size_t total = woff[n];
    const size_t o_bits = 0, o_len = o_bits + (((size_t)n * bits_stride + 15) & ~(size_t)15), o_data = o_len + (size_t)n * 4,
                 o_sig = o_data + (size_t)n * 128, o_st = o_sig + (size_t)n * 96, o_end = o_st + (size_t)n * 4;
    int rc;
    if ((rc = ensure(ctx, ctx->in_a, total + 16)) || (rc = ensure(ctx, ctx->in_b, (size_t)(n + 1) * 4)) || (rc = ensure(ctx, ctx->out_a, o_end))) return rc;
    if (total) CK(cudaMemcpyAsync(ctx->in_a.p, wire, total, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_b.p, woff, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, s));
    uint8_t* o = (uint8_t*)ctx->out_a.p;
    k_attestations_decode<<<blocks_for((uint64_t)n * 32, 128), 128, 0, s>>>((const uint8_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, n, bits_stride, max_bits,
                                                                          o + o_bits, (uint32_t*)(o + o_len), o + o_data, o + o_sig, (int32_t*)(o + o_st));
    CKL(ctx);
    std::vector<uint8_t> host(o_end);                       // outputs stay untouched on error
    CK(cudaMemcpyAsync(host.data(), o, o_end, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    memcpy(bits_out, host.data() + o_bits, (size_t)n * bits_stride);
    memcpy(bit_len_out, host.data() + o_len, (size_t)n * 4);
    memcpy(data128_out, host.data() + o_data, (size_t)n * 128);
    memcpy(sig96_out, host.data() + o_sig, (size_t)n * 96);
    memcpy(status_out, host.data() + o_st, (size_t)n * 4);
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ committee shuffle
int b2_shuffle_committees_dev(b2_ctx* ctx, const uint8_t* d_seed32, const uint32_t* d_active, uint32_t n_active, uint32_t rounds,
                              uint32_t* d_members_out, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_seed32 && d_members_out && rounds <= 255, "shuffle_committees_dev: bad arguments");
    if (n_active == 0) return B2_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const uint32_t nblk = (n_active + 255) / 256;
    int rc;
    if ((rc = ensure(ctx, ctx->sc_shuf, (size_t)rounds * nblk * 32 + 16)) || (rc = ensure(ctx, ctx->sc_pivot, (size_t)rounds * 8 + 16))) return rc;
    if (rounds) {
        k_shuffle_sources<<<blocks_for((uint64_t)rounds * (nblk + 1), 128), 128, 0, s>>>(d_seed32, n_active, rounds, nblk, (uint8_t*)ctx->sc_shuf.p,
                                                                                  (unsigned long long*)ctx->sc_pivot.p);
        CKL(ctx);
    }
    k_shuffle_apply<<<blocks_for(n_active, 256), 256, 0, s>>>(n_active, rounds, nblk, (const uint8_t*)ctx->sc_shuf.p,
                                                             (const unsigned long long*)ctx->sc_pivot.p, d_active, d_members_out);
    CKL(ctx);
    return B2_OK;
}

int b2_shuffle_committees(b2_ctx* ctx, const uint8_t* seed32, const uint32_t* active, uint32_t n_active, uint32_t rounds, uint32_t* members_out) {
    B2_NVTX;
    REQUIRE(ctx && seed32 && members_out && rounds <= 255, "shuffle_committees: bad arguments");
    if (n_active == 0) return B2_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_e, 32)) || (rc = ensure(ctx, ctx->in_a, (size_t)n_active * 4)) || (rc = ensure(ctx, ctx->out_a, (size_t)n_active * 4)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_e.p, seed32, 32, cudaMemcpyHostToDevice, s));
    if (active) CK(cudaMemcpyAsync(ctx->in_a.p, active, (size_t)n_active * 4, cudaMemcpyHostToDevice, s));
    if ((rc = b2_shuffle_committees_dev(ctx, (const uint8_t*)ctx->in_e.p, active ? (const uint32_t*)ctx->in_a.p : nullptr, n_active, rounds,
                                        (uint32_t*)ctx->out_a.p, s)))
        return rc;
    CK(cudaMemcpyAsync(members_out, ctx->out_a.p, (size_t)n_active * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ latest messages
int b2_latest_messages_reset(b2_ctx* ctx) {
    B2_NVTX;
    REQUIRE(ctx && ctx->n_val > 0, "latest_messages_reset: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemsetAsync(ctx->d_lmd_key, 0, ctx->n_val * 8, ctx->s_main));
    CK(cudaMemsetAsync(ctx->d_lmd_block, 0, ctx->n_val * 4, ctx->s_main));
    CK(cudaMemsetAsync(ctx->d_equiv, 0, ctx->n_val, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return B2_OK;
}

int b2_latest_messages_load(b2_ctx* ctx, const uint64_t* epoch, const uint32_t* block_idx, const uint8_t* has_msg, const uint8_t* equivocating,
                            uint64_t n) {
    B2_NVTX;
    REQUIRE(ctx && epoch && block_idx && has_msg && equivocating && n == ctx->n_val && n > 0, "latest_messages_load: bad arguments / registry size");
    std::vector<unsigned long long> key(n);
    for (uint64_t v = 0; v < n; v++) {
        REQUIRE(!has_msg[v] || epoch[v] < 0xffffffffull, "latest_messages_load: epoch does not fit 32 bits");
        key[v] = has_msg[v] ? ((epoch[v] << 32) | 0xffffffffull) : 0ull;
    }
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpyAsync(ctx->d_lmd_key, key.data(), n * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_lmd_block, block_idx, n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_equiv, equivocating, n, cudaMemcpyHostToDevice, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

int b2_latest_messages_read(b2_ctx* ctx, uint64_t* epoch, uint32_t* block_idx, uint8_t* has_msg, uint64_t n) {
    B2_NVTX;
    REQUIRE(ctx && epoch && block_idx && has_msg && n == ctx->n_val && n > 0, "latest_messages_read: bad arguments / registry size");
    std::vector<unsigned long long> key(n);
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpyAsync(key.data(), ctx->d_lmd_key, n * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaMemcpyAsync(block_idx, ctx->d_lmd_block, n * 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    for (uint64_t v = 0; v < n; v++) {
        has_msg[v] = key[v] != 0;
        epoch[v] = key[v] >> 32;
        if (!has_msg[v]) block_idx[v] = 0;
    }
    return B2_OK;
}

int b2_latest_messages_update_dev(b2_ctx* ctx, const uint32_t* d_members, const uint32_t* d_off, const uint8_t* d_bits, uint32_t bits_stride,
                                  const uint64_t* d_target_epoch, const uint32_t* d_block_idx, const uint8_t* d_accept, uint32_t n_agg, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_members && d_off && d_bits && d_target_epoch && d_block_idx && n_agg > 0 && bits_stride > 0, "latest_messages_update_dev: bad arguments");
    REQUIRE(ctx->n_val > 0, "latest_messages_update: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    k_lmd_phase1<<<n_agg, 128, 0, s>>>(d_members, d_off, d_bits, bits_stride, d_target_epoch, d_accept, ctx->d_equiv, n_agg, ctx->d_lmd_key,
                                       ctx->n_val, ctx->d_guard);
    CKL(ctx);
    k_lmd_phase2<<<n_agg, 128, 0, s>>>(d_members, d_off, d_bits, bits_stride, d_target_epoch, d_block_idx, d_accept, ctx->d_equiv, n_agg,
                                       ctx->d_lmd_key, ctx->d_lmd_block, ctx->n_val);
    CKL(ctx);
    return B2_OK;
}

int b2_latest_messages_update(b2_ctx* ctx, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                              const uint64_t* target_epoch, const uint32_t* block_idx, const uint8_t* accept, uint32_t n_agg) {
    B2_NVTX;
    REQUIRE(ctx && members && off && bits && target_epoch && block_idx && n_agg > 0 && bits_stride > 0, "latest_messages_update: bad arguments");
    REQUIRE(ctx->n_val > 0, "latest_messages_update: registry not loaded");
    int rc;
    if ((rc = check_batch(ctx, members, off, bits_stride, n_agg, ctx->n_val))) return rc;
    for (uint32_t a = 0; a < n_agg; a++) REQUIRE(target_epoch[a] < 0xffffffffull, "latest_messages_update: epoch does not fit 32 bits");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    if ((rc = upload_batch(ctx, members, off, bits, bits_stride, n_agg, s))) return rc;
    if ((rc = ensure(ctx, ctx->in_d, (size_t)n_agg * 8)) || (rc = ensure(ctx, ctx->in_e, (size_t)n_agg * 4)) || (rc = ensure(ctx, ctx->in_f, n_agg)))
        return rc;
    CK(cudaMemcpyAsync(ctx->in_d.p, target_epoch, (size_t)n_agg * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_e.p, block_idx, (size_t)n_agg * 4, cudaMemcpyHostToDevice, s));
    if (accept) CK(cudaMemcpyAsync(ctx->in_f.p, accept, n_agg, cudaMemcpyHostToDevice, s));
    if ((rc = b2_latest_messages_update_dev(ctx, (const uint32_t*)ctx->in_a.p, (const uint32_t*)ctx->in_b.p, (const uint8_t*)ctx->in_c.p, bits_stride,
                                            (const uint64_t*)ctx->in_d.p, (const uint32_t*)ctx->in_e.p, accept ? (const uint8_t*)ctx->in_f.p : nullptr,
                                            n_agg, s)))
        return rc;
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ participation flags (process_attestation :738-754)
int b2_participation_load(b2_ctx* ctx, int which, const uint8_t* participation, uint64_t n) {
    B2_NVTX;
    REQUIRE(ctx && (which == 0 || which == 1) && participation && n == ctx->n_val && n > 0, "participation_load: bad arguments / registry size");
    CK(cudaSetDevice(ctx->device));
    int rc;
    const size_t words = (n + 3) / 4;
    if (!ctx->d_part[which] && (rc = dev_alloc(ctx, &ctx->d_part[which], words))) return rc;
    if (!ctx->d_part_first) {
        if ((rc = dev_alloc(ctx, &ctx->d_part_first, 3 * n))) return rc;
        CK(cudaMemsetAsync(ctx->d_part_first, 0xff, 3 * n * 4, ctx->s_main));
    }
    CK(cudaMemsetAsync(ctx->d_part[which], 0, words * 4, ctx->s_main));
    CK(cudaMemcpyAsync(ctx->d_part[which], participation, n, cudaMemcpyHostToDevice, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return B2_OK;
}
int b2_participation_read(b2_ctx* ctx, int which, uint8_t* participation_out, uint64_t n) {
    B2_NVTX;
    REQUIRE(ctx && (which == 0 || which == 1) && participation_out && n == ctx->n_val && ctx->d_part[which], "participation_read: not loaded / bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(participation_out, ctx->d_part[which], n, cudaMemcpyDeviceToHost, ctx->s_main));
    CK(cudaStreamSynchronize(ctx->s_main));
    return B2_OK;
}
int b2_participation_update(b2_ctx* ctx, int which, const uint32_t* members, const uint32_t* off, const uint8_t* bits, uint32_t bits_stride,
                            const uint8_t* flag_mask, const uint8_t* accept, uint32_t n_agg, uint64_t effective_balance_increment,
                            uint64_t base_reward_per_increment, uint64_t* numerator_out) {
    B2_NVTX;
    REQUIRE(ctx && (which == 0 || which == 1) && members && off && bits && flag_mask && numerator_out && n_agg > 0 && bits_stride > 0 &&
                effective_balance_increment > 0, "participation_update: bad arguments");
    REQUIRE(ctx->n_val > 0 && ctx->d_part[which], "participation_update: registry / participation table not loaded");
    int rc;
    if ((rc = check_batch(ctx, members, off, bits_stride, n_agg, ctx->n_val))) return rc;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    if ((rc = upload_batch(ctx, members, off, bits, bits_stride, n_agg, s))) return rc;
    if ((rc = ensure(ctx, ctx->in_d, n_agg)) || (rc = ensure(ctx, ctx->in_f, n_agg)) || (rc = ensure(ctx, ctx->out_a, (size_t)n_agg * 8))) return rc;
    CK(cudaMemcpyAsync(ctx->in_d.p, flag_mask, n_agg, cudaMemcpyHostToDevice, s));
    if (accept) CK(cudaMemcpyAsync(ctx->in_f.p, accept, n_agg, cudaMemcpyHostToDevice, s));
    const uint32_t *dm = (const uint32_t*)ctx->in_a.p, *dof = (const uint32_t*)ctx->in_b.p;
    const uint8_t *db = (const uint8_t*)ctx->in_c.p, *dmask = (const uint8_t*)ctx->in_d.p, *dacc = accept ? (const uint8_t*)ctx->in_f.p : nullptr;
    k_part_phase1<<<n_agg, 128, 0, s>>>(dm, dof, db, bits_stride, dmask, dacc, n_agg, (const uint8_t*)ctx->d_part[which], ctx->d_part_first);
    CKL(ctx);
    k_part_phase2<<<n_agg, 128, 0, s>>>(dm, dof, db, bits_stride, dmask, dacc, n_agg, ctx->d_part_first, ctx->d_eff, effective_balance_increment,
                                        base_reward_per_increment, (unsigned long long*)ctx->out_a.p);
    CKL(ctx);
    k_part_phase3<<<n_agg, 128, 0, s>>>(dm, dof, db, bits_stride, dmask, dacc, n_agg, ctx->d_part_first, ctx->d_part[which]);
    CKL(ctx);
    CK(cudaMemcpyAsync(numerator_out, ctx->out_a.p, (size_t)n_agg * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ FFG balance sums (:793-803)
int b2_ffg_balances(b2_ctx* ctx, uint32_t flag_index, uint64_t* out4) {
    B2_NVTX;
    REQUIRE(ctx && out4 && flag_index < 8, "ffg_balances: bad arguments");
    REQUIRE(ctx->n_val > 0, "ffg_balances: registry not loaded");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->out_b, 32))) return rc;
    CK(cudaMemsetAsync(ctx->out_b.p, 0, 32, s));
    k_ffg_balances<<<ctx->n_sm * 4, 256, 0, s>>>(ctx->n_val, (const unsigned long long*)ctx->d_eff, ctx->d_flags, (const uint8_t*)ctx->d_part[0],
                                                 (const uint8_t*)ctx->d_part[1], flag_index, (unsigned long long*)ctx->out_b.p);
    CKL(ctx);
    CK(cudaMemcpyAsync(out4, ctx->out_b.p, 32, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ fork-choice variants
int b2_set_fork_choice_params(b2_ctx* ctx, uint64_t min_vote_epoch, int exclude_slashed) {
    B2_NVTX;
    REQUIRE(ctx && min_vote_epoch < 0xffffffffull, "set_fork_choice_params: bad arguments");
    ctx->fc_min_key = (unsigned long long)min_vote_epoch << 32;
    ctx->fc_exclude_slashed = exclude_slashed ? 1 : 0;
    return B2_OK;
}

int b2_set_verify_mode(b2_ctx* ctx, int mode, const uint8_t* seed32) {
    B2_NVTX;
    REQUIRE(ctx && (mode == 0 || (mode == 1 && seed32)), "set_verify_mode: bad arguments (mode 1 needs a 32-byte seed)");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());                   // no epoch in flight may see the mode change half way
    if (mode == 1) CK(cudaMemcpy(ctx->d_rlc_seed, seed32, 32, cudaMemcpyHostToDevice));
    ctx->verify_mode = mode;
    return B2_OK;
}

int b2_on_attester_slashing(b2_ctx* ctx, const uint32_t* indices_1, uint32_t n1, const uint32_t* indices_2, uint32_t n2) {
    B2_NVTX;
    REQUIRE(ctx && ctx->n_val > 0 && (n1 == 0 || indices_1) && (n2 == 0 || indices_2), "on_attester_slashing: bad arguments / registry not loaded");
    for (uint32_t i = 1; i < n1; i++) REQUIRE(indices_1[i - 1] < indices_1[i], "on_attester_slashing: attesting_indices must be sorted and unique");
    for (uint32_t i = 1; i < n2; i++) REQUIRE(indices_2[i - 1] < indices_2[i], "on_attester_slashing: attesting_indices must be sorted and unique");
    if (n1 == 0 || n2 == 0) return B2_OK;
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    int rc;
    if ((rc = ensure(ctx, ctx->in_a, (size_t)n1 * 4)) || (rc = ensure(ctx, ctx->in_b, (size_t)n2 * 4))) return rc;
    CK(cudaMemcpyAsync(ctx->in_a.p, indices_1, (size_t)n1 * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->in_b.p, indices_2, (size_t)n2 * 4, cudaMemcpyHostToDevice, s));
    k_mark_equivocating<<<blocks_for(n1, 128), 128, 0, s>>>((const uint32_t*)ctx->in_a.p, n1, (const uint32_t*)ctx->in_b.p, n2, ctx->n_val, ctx->d_equiv);
    CKL(ctx);
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

// ------------------------------------------------------------------------------------------ block tree
int b2_tree_load(b2_ctx* ctx, const uint32_t* parent, const uint64_t* slot, const uint8_t* root32, const uint8_t* leaf_viable, uint32_t n) {
    B2_NVTX;
    REQUIRE(ctx && parent && slot && root32 && leaf_viable && n > 0, "tree_load: bad arguments");
    for (uint32_t b = 1; b < n; b++) REQUIRE(parent[b] < b, "tree_load: blocks must be in topological order (parent[b] < b)");
    // children in CSR form
    std::vector<uint32_t> child_off(n + 1, 0), child_idx(n > 1 ? n - 1 : 0), fill(n, 0);
    for (uint32_t b = 1; b < n; b++) child_off[parent[b] + 1]++;
    for (uint32_t b = 0; b < n; b++) child_off[b + 1] += child_off[b];
    for (uint32_t b = 1; b < n; b++) child_idx[child_off[parent[b]] + fill[parent[b]]++] = b;
    // DFS pre-order numbering and subtree sizes
    std::vector<uint32_t> pre(n), size(n, 1), stack;
    stack.reserve(n);
    stack.push_back(0);
    uint32_t counter = 0;
    while (!stack.empty()) {
        uint32_t b = stack.back();
        stack.pop_back();
        pre[b] = counter++;
        for (uint32_t k = child_off[b + 1]; k > child_off[b]; k--) stack.push_back(child_idx[k - 1]);
    }
    for (uint32_t b = n - 1; b >= 1; b--) size[parent[b]] += size[b];
    // get_filtered_block_tree: keep a block iff a viable leaf lies below it
    std::vector<uint8_t> keep(n, 0);
    for (uint32_t b = 0; b < n; b++)
        if (child_off[b + 1] == child_off[b]) keep[b] = leaf_viable[b] ? 1 : 0;
    for (uint32_t b = n - 1; b >= 1; b--)
        if (keep[b]) keep[parent[b]] = 1;
    // lexicographic rank of the roots (the tie-break of get_head compares 32-byte roots)
    std::vector<uint32_t> order(n), rank(n);
    for (uint32_t b = 0; b < n; b++) order[b] = b;
    std::sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
        int c = memcmp(root32 + 32ull * x, root32 + 32ull * y, 32);
        return c != 0 ? c < 0 : x < y;
    });
    for (uint32_t i = 0; i < n; i++) rank[order[i]] = i;
    CK(cudaSetDevice(ctx->device));
    int rc;
    // everything the tree kernel reads is stored in pre-order
    std::vector<uint32_t> inv(n), size_keep(n), rank_p(n), packed(n);
    for (uint32_t b = 0; b < n; b++) {
        const uint32_t p = pre[b];
        inv[p] = b;
        size_keep[p] = size[b] | (keep[b] ? 0x80000000u : 0u);
        rank_p[p] = rank[b];
        packed[p] = (size[b] & 0x7fffu) | ((rank[b] & 0x7fffu) << 15) | (keep[b] ? 0x80000000u : 0u);
    }
    if ((rc = dev_alloc(ctx, &ctx->d_pre, n)) || (rc = dev_alloc(ctx, &ctx->d_inv, n)) || (rc = dev_alloc(ctx, &ctx->d_size_keep, n)) ||
        (rc = dev_alloc(ctx, &ctx->d_rank, n)) || (rc = dev_alloc(ctx, &ctx->d_next, (size_t)n + 1)) || (rc = dev_alloc(ctx, &ctx->d_gsize, n)) ||
        (rc = dev_alloc(ctx, &ctx->d_votes, n)) || (rc = dev_alloc(ctx, &ctx->d_prefix, (size_t)n + 1)) || (rc = dev_alloc(ctx, &ctx->d_w2, n)) ||
        (rc = dev_alloc(ctx, &ctx->d_weight, n)) || (rc = dev_alloc(ctx, &ctx->d_head, 1)) || (rc = dev_alloc(ctx, &ctx->d_packed, n)))
        return rc;
    cudaStream_t s = ctx->s_main;
    CK(cudaMemcpyAsync(ctx->d_packed, packed.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    ctx->packed_ok = n < 32768;
    CK(cudaMemcpyAsync(ctx->d_pre, pre.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_inv, inv.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_size_keep, size_keep.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(ctx->d_rank, rank_p.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(ctx->d_votes, 0, (size_t)n * 8, s));
    CK(cudaStreamSynchronize(s));
    ctx->n_blocks = n;
    return B2_OK;
}

int b2_vote_weights_range_dev(b2_ctx* ctx, uint64_t v_begin, uint64_t v_end, uint64_t* d_votes_preorder, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_votes_preorder, "vote_weights_dev: bad arguments");
    REQUIRE(ctx->n_val > 0 && ctx->n_blocks > 0, "vote_weights: registry or tree not loaded");
    REQUIRE(v_begin <= v_end && v_end <= ctx->n_val, "vote_weights_range: validator range outside the registry");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    const uint64_t n = v_end - v_begin;
    const size_t bins_bytes = (size_t)ctx->n_blocks * 8;
    if (n == 0) {
        // nothing to scatter; the accumulator is clean by contract
    } else if (bins_bytes <= 200 * 1024) {
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->n_sm, (n + 1023) / 1024);
        k_ghost_votes_smem<<<grid, 1024, bins_bytes, s>>>(n, ctx->d_lmd_key + v_begin, ctx->d_lmd_block + v_begin, ctx->d_equiv + v_begin, ctx->d_flags + v_begin,
                                                         ctx->d_eff + v_begin, ctx->d_pre, ctx->n_blocks, (unsigned long long*)d_votes_preorder, ctx->fc_min_key, 1u,
                                                         ctx->fc_exclude_slashed ? 3u : 1u);
        CKL(ctx);
    } else {
        k_ghost_votes<<<blocks_for(n, 256), 256, 0, s>>>(n, ctx->d_lmd_key + v_begin, ctx->d_lmd_block + v_begin, ctx->d_equiv + v_begin, ctx->d_flags + v_begin,
                                                         ctx->d_eff + v_begin, ctx->d_pre, ctx->n_blocks, (unsigned long long*)d_votes_preorder, ctx->fc_min_key, 1u,
                                                         ctx->fc_exclude_slashed ? 3u : 1u);
        CKL(ctx);
    }
    CK(cudaEventRecord(ctx->ev_votes_done, s));
    return B2_OK;
}
int b2_vote_weights_dev(b2_ctx* ctx, uint64_t* d_votes_preorder, void* stream) {
    B2_NVTX;
    REQUIRE(ctx, "vote_weights_dev: bad arguments");
    return b2_vote_weights_range_dev(ctx, 0, ctx->n_val, d_votes_preorder, stream);
}

int b2_head_from_votes_dev(b2_ctx* ctx, uint64_t* d_votes_preorder, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score,
                           uint64_t* d_weight_out, uint32_t* d_head_idx_out, void* stream) {
    B2_NVTX;
    REQUIRE(ctx && d_votes_preorder && d_head_idx_out, "head_from_votes_dev: bad arguments");
    REQUIRE(ctx->n_blocks > 0, "head_from_votes: tree not loaded");
    REQUIRE(justified_idx < ctx->n_blocks && boost_idx < (int32_t)ctx->n_blocks, "head_from_votes: block index out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = (cudaStream_t)stream;
    ghost_tree_args A;
    A.n = ctx->n_blocks;
    A.pre = ctx->d_pre;
    A.inv = ctx->d_inv;
    A.size_keep = ctx->d_size_keep;
    A.rank = ctx->d_rank;
    A.packed = ctx->packed_ok ? ctx->d_packed : nullptr;
    A.votes = (unsigned long long*)d_votes_preorder;
    A.g_w = ctx->d_prefix;
    A.g_w2 = ctx->d_w2;
    A.g_size = ctx->d_gsize;
    A.g_next = ctx->d_next;
    A.weight_out = (unsigned long long*)d_weight_out;
    A.head_out = d_head_idx_out;
    A.justified = justified_idx;
    A.boost_idx = boost_idx;
    A.boost_score = boost_score;
    size_t smem = ((size_t)A.n + 1) * 8 + (size_t)A.n * 8 + 8;
    A.use_smem = smem <= 226 * 1024 && A.n <= 15 * 1024;   // 227 KB per block minus the kernel's static shared memory
    A.dbg = ctx->d_dbg;
    A.hard_list = 0;
    if (A.use_smem && A.packed) smem = tree_smem_bytes(A);
    k_ghost_tree<<<1, 1024, A.use_smem ? smem : 0, s>>>(A);
    CKL(ctx);
    return B2_OK;
}

int b2_get_weights(b2_ctx* ctx, int32_t boost_idx, uint64_t boost_score, uint64_t* weight_out) {
    B2_NVTX;
    REQUIRE(ctx && weight_out, "get_weights: bad arguments");
    REQUIRE(ctx->n_val > 0 && ctx->n_blocks > 0, "get_weights: registry or tree not loaded");
    int rc;
    cudaStream_t s = ctx->s_main;
    if ((rc = b2_vote_weights_dev(ctx, (uint64_t*)ctx->d_votes, s))) return rc;
    if ((rc = b2_head_from_votes_dev(ctx, (uint64_t*)ctx->d_votes, 0, boost_idx, boost_score, (uint64_t*)ctx->d_weight, ctx->d_head, s))) return rc;
    CK(cudaMemcpyAsync(weight_out, ctx->d_weight, (size_t)ctx->n_blocks * 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    return B2_OK;
}

static void fill_tree_args(b2_ctx* ctx, ghost_tree_args& A, uint64_t* d_votes_preorder, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score,
                           uint64_t* d_weight_out, uint32_t* d_head_idx_out);

// ---- multi-GPU get_head with the all-reduce fused in over NVLink peer memory (k_get_head_fused_nvl)
static size_t fc_acc_bytes(uint32_t n_blocks) { return ((size_t)2 * n_blocks * 8 + 255) & ~(size_t)255; }
int b2_fc_exchange_export(b2_ctx* ctx, uint8_t* handle64_out) {
    B2_NVTX;
    REQUIRE(ctx && handle64_out, "fc_exchange_export: bad arguments");
    REQUIRE(ctx->n_blocks > 0, "fc_exchange_export: load the block tree first (the exchange block is sized by it)");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle is 64 bytes");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    for (int r = 0; r < B2_MAX_PEERS; r++) {
        if (ctx->fc_peer[r] && r != ctx->fc_rank) CK(cudaIpcCloseMemHandle(ctx->fc_peer[r]));
        ctx->fc_peer[r] = nullptr;
    }
    if (ctx->fc_block) CK(cudaFree(ctx->fc_block));
    ctx->fc_block = nullptr;
    ctx->fc_rank = -1;
    ctx->fc_world = 0;
    ctx->fc_block_bytes = fc_acc_bytes(ctx->n_blocks) + 2 * B2_MAX_PEERS * sizeof(unsigned int);
    CK(cudaMalloc(&ctx->fc_block, ctx->fc_block_bytes));
    CK(cudaMemset(ctx->fc_block, 0, ctx->fc_block_bytes));
    ctx->fc_blocks_cap = ctx->n_blocks;
    ctx->fc_seq = 0;
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, ctx->fc_block));
    memcpy(handle64_out, &h, 64);
    return B2_OK;
}
int b2_fc_exchange_open(b2_ctx* ctx, int rank, int world, const uint8_t* handles64) {
    B2_NVTX;
    REQUIRE(ctx && handles64 && world >= 1 && world <= B2_MAX_PEERS && rank >= 0 && rank < world, "fc_exchange_open: bad arguments");
    REQUIRE(ctx->fc_block, "fc_exchange_open: call b2_fc_exchange_export first");
    CK(cudaSetDevice(ctx->device));
    for (int r = 0; r < world; r++) {
        if (r == rank) {
            ctx->fc_peer[r] = ctx->fc_block;
            continue;
        }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles64 + 64 * (size_t)r, 64);
        CK(cudaIpcOpenMemHandle(&ctx->fc_peer[r], h, cudaIpcMemLazyEnablePeerAccess));
    }
    ctx->fc_rank = rank;
    ctx->fc_world = world;
    return B2_OK;
}
int b2_get_head_multi(b2_ctx* ctx, uint64_t v_begin, uint64_t v_end, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score, uint32_t* head_idx_out) {
    B2_NVTX;
    REQUIRE(ctx && head_idx_out, "get_head_multi: bad arguments");
    REQUIRE(ctx->n_val > 0 && ctx->n_blocks > 0, "get_head_multi: registry or tree not loaded");
    REQUIRE(ctx->fc_world >= 1 && ctx->fc_blocks_cap == ctx->n_blocks, "get_head_multi: exchange block not opened for this tree (b2_fc_exchange_export / _open)");
    REQUIRE(v_begin <= v_end && v_end <= ctx->n_val, "get_head_multi: validator range outside the registry");
    REQUIRE(justified_idx < ctx->n_blocks && boost_idx < (int32_t)ctx->n_blocks, "get_head_multi: block index out of range");
    CK(cudaSetDevice(ctx->device));
    cudaStream_t s = ctx->s_main;
    ghost_tree_args A;
    fill_tree_args(ctx, A, (uint64_t*)ctx->d_votes, justified_idx, boost_idx, boost_score, nullptr, ctx->d_head);
    REQUIRE(A.use_smem && A.packed, "get_head_multi: the tree does not fit the shared-memory form (< ~14 500 blocks)");
    const uint64_t n = v_end - v_begin;
    ghost_votes_args V = {n, ctx->d_lmd_key + v_begin, ctx->d_lmd_block + v_begin, ctx->d_equiv + v_begin, ctx->d_flags + v_begin, ctx->d_eff + v_begin,
                          ctx->fc_min_key, 1u, ctx->fc_exclude_slashed ? 3u : 1u};
    ghost_peer_args P;
    const uint32_t seq = ++ctx->fc_seq ? ctx->fc_seq : ++ctx->fc_seq;                  // the collective call number (same on every rank)
    const uint32_t hseq = ++ctx->head_seq ? ctx->head_seq : ++ctx->head_seq;            // this context's host-slot sequence number
    for (int r = 0; r < B2_MAX_PEERS; r++) {
        P.acc[r] = r < ctx->fc_world ? (unsigned long long*)ctx->fc_peer[r] : nullptr;
        P.flags[r] = r < ctx->fc_world ? (unsigned int*)((char*)ctx->fc_peer[r] + fc_acc_bytes(ctx->n_blocks)) : nullptr;
    }
    P.rank = (uint32_t)ctx->fc_rank;
    P.world = (uint32_t)ctx->fc_world;
    P.seq = seq;
    const size_t smem = tree_smem_bytes(A);
    const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>((uint64_t)ctx->n_sm, (n + 1023) / 1024));
    k_get_head_fused_nvl<<<grid, 1024, smem, s>>>(V, A, P, ctx->d_ticket, (unsigned long long*)ctx->d_votes, ctx->d_head_host, hseq);
    CKL(ctx);
    CK(cudaEventRecord(ctx->ev_votes_done, s));
    volatile unsigned long long* slot64 = reinterpret_cast<volatile unsigned long long*>(ctx->h_head);
    unsigned long long got = 0;
    for (uint64_t spins = 0; (uint32_t)((got = *slot64) >> 32) != hseq; spins++) {
        if ((spins & 0xfffff) == 0xfffff) {
            cudaError_t q = cudaStreamQuery(s);
            if (q != cudaErrorNotReady && q != cudaSuccess) return fail_cuda(ctx, q, "get_head_multi kernel");
        }
    }
    *head_idx_out = (uint32_t)got;
    return B2_OK;
}

static void fill_tree_args(b2_ctx* ctx, ghost_tree_args& A, uint64_t* d_votes_preorder, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score,
                           uint64_t* d_weight_out, uint32_t* d_head_idx_out) {
    A.n = ctx->n_blocks;
    A.pre = ctx->d_pre;
    A.inv = ctx->d_inv;
    A.size_keep = ctx->d_size_keep;
    A.rank = ctx->d_rank;
    A.packed = ctx->packed_ok ? ctx->d_packed : nullptr;
    A.votes = (unsigned long long*)d_votes_preorder;
    A.g_w = ctx->d_prefix;
    A.g_w2 = ctx->d_w2;
    A.g_size = ctx->d_gsize;
    A.g_next = ctx->d_next;
    A.weight_out = (unsigned long long*)d_weight_out;
    A.head_out = d_head_idx_out;
    A.justified = justified_idx;
    A.boost_idx = boost_idx;
    A.boost_score = boost_score;
    const size_t smem = ((size_t)A.n + 1) * 8 + (size_t)A.n * 8 + 8;
    A.use_smem = smem <= 226 * 1024 && A.n <= 15 * 1024;   // 227 KB per block minus the kernel's static shared memory
    A.dbg = ctx->d_dbg;
    A.hard_list = 0;
}

// Diagnostics: SM-clock stamps (clock64) of the phases of the LAST b2_get_head / b2_head_from_votes_dev of this context: out32[0..7] =
// tree phase boundaries (start, staged, scanned, weights, weights stored, marked, counted, head found), out32[15] = tree end,
// out32[16..19] = vote scatter of CTA 0 (start, bins zeroed, scattered, flushed).  For profiling (tools/, profiles/), not for results.
int b2_debug_head_clocks(b2_ctx* ctx, uint64_t* out32) {
    REQUIRE(ctx && out32, "debug_head_clocks: bad arguments");
    CK(cudaSetDevice(ctx->device));
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out32, ctx->d_dbg, 32 * 8, cudaMemcpyDeviceToHost));
    return B2_OK;
}

int b2_get_head(b2_ctx* ctx, uint32_t justified_idx, int32_t boost_idx, uint64_t boost_score, uint32_t* head_idx_out) {
    B2_NVTX;
    REQUIRE(ctx && head_idx_out, "get_head: bad arguments");
    REQUIRE(ctx->n_val > 0 && ctx->n_blocks > 0, "get_head: registry or tree not loaded");
    REQUIRE(justified_idx < ctx->n_blocks && boost_idx < (int32_t)ctx->n_blocks, "get_head: block index out of range");
    int rc;
    cudaStream_t s = ctx->s_main;
    ghost_tree_args A;
    fill_tree_args(ctx, A, (uint64_t*)ctx->d_votes, justified_idx, boost_idx, boost_score, nullptr, ctx->d_head);
    if (ctx->head_fused && A.use_smem && A.packed) {
        // one launch; the kernel writes (head, sequence) into mapped pinned memory and the host spins on the sequence number
        CK(cudaSetDevice(ctx->device));
        ghost_votes_args V = {ctx->n_val, ctx->d_lmd_key, ctx->d_lmd_block, ctx->d_equiv, ctx->d_flags, ctx->d_eff, ctx->fc_min_key, 1u,
                              ctx->fc_exclude_slashed ? 3u : 1u};
        const uint32_t seq = ++ctx->head_seq ? ctx->head_seq : ++ctx->head_seq;       // never 0
        const size_t smem = tree_smem_bytes(A);
        const unsigned grid = (unsigned)std::min<uint64_t>((uint64_t)ctx->n_sm, (ctx->n_val + 1023) / 1024);
        k_get_head_fused<<<grid, 1024, smem, s>>>(V, A, ctx->d_ticket, ctx->d_head_host, seq);
        CKL(ctx);
        CK(cudaEventRecord(ctx->ev_votes_done, s));
        volatile unsigned long long* slot64 = reinterpret_cast<volatile unsigned long long*>(ctx->h_head);
        unsigned long long got = 0;
        for (uint64_t spins = 0; (uint32_t)((got = *slot64) >> 32) != seq; spins++) {
            if ((spins & 0xfffff) == 0xfffff) {                 // every ~1M polls: has the kernel died?
                cudaError_t q = cudaStreamQuery(s);
                if (q != cudaErrorNotReady && q != cudaSuccess) return fail_cuda(ctx, q, "get_head kernel");
                if (q == cudaSuccess && (uint32_t)(*slot64 >> 32) != seq) {    // finished without publishing: should not happen; fall back to a copy
                    uint32_t h = 0;
                    CK(cudaMemcpy(&h, ctx->d_head, 4, cudaMemcpyDeviceToHost));
                    *head_idx_out = h;
                    return B2_OK;
                }
            }
        }
        *head_idx_out = (uint32_t)got;
        return B2_OK;
    }
    if ((rc = b2_vote_weights_dev(ctx, (uint64_t*)ctx->d_votes, s))) return rc;
    if ((rc = b2_head_from_votes_dev(ctx, (uint64_t*)ctx->d_votes, justified_idx, boost_idx, boost_score, nullptr, ctx->d_head, s))) return rc;
    uint32_t h = 0;
    CK(cudaMemcpyAsync(&h, ctx->d_head, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    *head_idx_out = h;
    return B2_OK;
}

}  // extern "C"
