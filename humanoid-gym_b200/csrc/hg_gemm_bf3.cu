// tcgen05 / TMEM / TMA GEMM on PRE-SPLIT bf16 operands ("bf16x3"), sm_100a.
//
//     x ~= x_hi + x_lo,   x_hi = bf16(x),   x_lo = bf16(x - x_hi)            (16 significant bits)
//     D += A_lo B_hi + A_hi B_lo + A_hi B_hi                                   three kind::f16 MMAs, fp32 accumulation in TMEM
//
// = 1.5 TF32-equivalent tensor passes (3xTF32 needs 3), relative error ~5e-6 per product, ~9e-6 on the
// ActorCritic gradients (bar: 1e-4).  Every operand arrives already split -- by the kernel that produced it
// (previous GEMM epilogue, minibatch gather, dgrad of the output head, the post-Adam weight split) -- as two bf16
// planes [2][rows][ld], so this kernel has NO splitter warps: TMA loads the four tiles {A_hi, A_lo, B_hi, B_lo}
// of a 64-k stage straight into 128B-swizzled shared memory and the tensor core reads them from there.
//
// One kernel covers the three products of a Linear layer (algo/ppo/actor_critic.py:54-77 and its autograd),
// without materialising any transpose:
//     forward  H  = ELU(X W^T + b)   A = X  (K-major)    B = W  (K-major)     epilogue: +bias, ELU, split store
//     dgrad    dZ' = (dZ W) * ELU'   A = dZ (K-major)    B = W  (MN-major)    epilogue: * ELU'(H), split store, column sums (bias grad)
//     wgrad    dW = dZ^T X           A = dZ (MN-major)   B = X  (MN-major)    split-K over the batch, fp32 atomics
//
// CTA = 128 x BN output tile (BN <= 256), persistent (one CTA per SM), 10 warps:
//     warps 0-7  epilogue     tcgen05.ld (warp w reads TMEM lanes 32*(w%4).., column half w/4) -> bias/ELU/ELU'/split/atomics
//     warp  8    TMA producer cp.async.bulk.tensor.3d ({k, rows, plane} boxes; OOB rows / k are zero-filled)
//     warp  9    MMA issuer   one lane issues 12 tcgen05.mma.kind::f16 (M=128, N=BN, K=16) per stage
// mbarrier ring of S stages (S = what fits in 227 KB), accumulator double-buffered in TMEM (2 x 256 columns) so the
// epilogue of tile i overlaps the main loop of tile i+1.
#include <cuda.h>
#include <stdlib.h>

#include <mutex>
#include <unordered_map>

#include "hg_common.cuh"
#include "hg_tc_ptx.cuh"

using namespace hgtc;

namespace {

constexpr int BM = 128, BK = 64;                     // 64 bf16 = one 128-byte swizzled row
constexpr int EPI_WARPS = 8;
constexpr int THREADS = (EPI_WARPS + 2) * 32;
constexpr int A_PLANE = BM * BK * 2;                 // 16 KB: one bf16 plane of the A tile
constexpr int A_BYTES = 2 * A_PLANE;                 // hi + lo
constexpr int SMEM_LIMIT = 227 * 1024;
constexpr int BAR_BYTES = 256;
constexpr int MAX_STAGES = 4;
constexpr int EPI_STAGE_WARP = 2 * 32 * 64;            // per epilogue warp: [2 planes][32 rows][32 bf16], SWIZZLE_64B
constexpr int EPI_STAGE_BYTES = 8 * EPI_STAGE_WARP;     // 32 KB

enum { EPI_F32 = 0, EPI_F32_BIAS = 1, EPI_SPLIT_BIAS_ELU = 2, EPI_SPLIT_DELU = 3, EPI_ATOMIC = 4, EPI_SPLIT = 5, EPI_DISCARD = 6 /* profiling: bias + ELU + split math, no stores */,
       EPI_SPLIT_DELU_TMA = 7 /* kernel-internal: EPI_SPLIT_DELU with the H blocks fetched by TMA (Args::h_tma) */ };

struct Args {
    float* C; int64_t ldc;
    uint16_t* Cs; int64_t ldcs, cs_plane;
    const float* bias;
    const uint16_t* Hs; int64_t ldhs, hs_plane;
    float* colsum;
    int M, N, K;
    int BN, a_mn, b_mn, epi;
    int kb_per_split, splits, stages;
    int h_tma;                           // dgrad: the H block of every 32 x 32 chunk arrives by TMA in a per-warp buffer (tmH valid)
    int cs_tma;                          // split outputs leave through shared memory + TMA stores (tmC valid)
    int relay;                           // pair form: 1 = each CTA's TMA signals its OWN barrier and a relay thread of the peer forwards one arrival per stage
    int dbg;                             // profiling (HG_BF3_DEBUG): 1 = no TMA loads (MMA on whatever is in smem), 2 = no MMAs (loads + commits only)
    long long* trace;                    // optional (hg_gemm_bf16x3_set_trace): [grid][8] cycle sums, see tools/bf3_trace.py
};

// shared-memory matrix descriptor (sm_100 UMMA, version 1), 16-bit operands, SWIZZLE_128B:
//   K-major  tile [rows][64 bf16]       : SBO = 1024 B (8 rows x 128 B); LBO unused
//   MN-major tile blocks of [64 k][64 mn]: SBO = 1024 B (8 k-rows x 128 B), LBO = byte distance between 64-wide MN blocks
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
    return d;
}
// instruction descriptor, kind::f16: D = F32, A = B = BF16, M = 128
__device__ __forceinline__ uint32_t make_idesc(int M, int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                                    // c_format F32
    d |= 1u << 7;                                    // a_format BF16
    d |= 1u << 10;                                   // b_format BF16
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(M >> 4) << 24;
    return d;
}

__device__ __forceinline__ unsigned long long hg_globaltimer() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
struct Work { int m0, n0, kb_begin, num_kb; };
// profiling: per-CTA event stamps of the first 64 k-blocks, behind the [grid][8] cycle sums: [grid][64][4] globaltimer ns
// (0 producer saw the stage free, 1 producer issued its loads, 2 peer relay saw its half land, 3 MMA thread saw the stage full)
#define BF3_STAMP(it_, slot_) do { if (g.trace && (it_) < 64) g.trace[(size_t)gridDim.x * 8 + ((size_t)blockIdx.x * 64 + (it_)) * 4 + (slot_)] = (long long)hg_globaltimer(); } while (0)
__device__ __forceinline__ Work decode_work(const Args& g, int w, int tiles_n, int tiles_mn, int num_kb_total, int tile_m = BM) {
    Work r;
    const int split = w / tiles_mn, t = w - split * tiles_mn;
    const int tm = t / tiles_n, tn = t - tm * tiles_n;
    r.m0 = tm * tile_m;
    r.n0 = tn * g.BN;
    r.kb_begin = split * g.kb_per_split;
    r.num_kb = min(num_kb_total, r.kb_begin + g.kb_per_split) - r.kb_begin;
    return r;
}

// (x0 at the lower address)
__device__ __forceinline__ uint32_t pack_bf16x2(float x0, float x1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(x1), "f"(x0));
    return r;
}
__device__ __forceinline__ float bf_lo(uint32_t p) { return __uint_as_float(p << 16); }
__device__ __forceinline__ float bf_hi(uint32_t p) { return __uint_as_float(p & 0xFFFF0000u); }

// nn.ELU(alpha=1) as exp(x) - 1 for x <= 0: ABSOLUTE error ~1e-7 (same form as torch's CUDA kernel).  The consumers are
// matrix products and ELU' = h + 1, for which only the absolute error matters; the epilogue is issue-bound, so the
// series that would restore relative accuracy near 0 is not spent here (the 3xTF32 rollout path keeps expm1).
__device__ __forceinline__ float elu1(float x) { return x > 0.0f ? x : exp_neg_fast(x) - 1.0f; }

// PAIR = true: the CTA-pair form (cluster of 2 on one TPC, tcgen05 cta_group::2).  The pair owns a 256 x BN tile: each CTA loads
// its own 128 rows of A and HALF of the B tile, the leader's single MMA thread issues M = 256 instructions that read B from
// both CTAs' shared memory, each CTA's TMEM receives its 128 rows and each CTA's epilogue warps drain them.  Per SM that halves
// the B bytes fetched over L2 -> SM and read from shared memory per MMA -- the two feeds that bound the single-CTA form
// (64 KB instead of 96 KB per 64-k stage at BN = 256, so three stages fit instead of two).
// EPI is a template parameter: one tight epilogue loop per instantiation (the all-in-one kernel was > 64 KB of SASS, most of it the
// other epilogues' branches, on a path where the 8 epilogue warps are the bottleneck of every short-K launch)
template <bool PAIR, int EPI>
__global__ void __launch_bounds__(THREADS, 1)
gemm_bf3_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmC,
                const __grid_constant__ CUtensorMap tmH, const Args g) {
    constexpr bool DELU = (EPI == EPI_SPLIT_DELU || EPI == EPI_SPLIT_DELU_TMA), HTMA = (EPI == EPI_SPLIT_DELU_TMA);
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;        // 0 = leader
    const int bn_cta = PAIR ? (g.BN >> 1) : g.BN;               // B rows this CTA holds
    const int b_plane = bn_cta * BK * 2;                        // bytes of one bf16 plane of this CTA's share of the B tile
    const int stage_bytes = A_BYTES + 2 * b_plane;
    unsigned char* epi_stage = smem + g.stages * stage_bytes;                         // 1024-aligned (stage_bytes is a multiple of 1 KB)
    unsigned char* h_stage = epi_stage + (g.cs_tma ? EPI_STAGE_BYTES : 0);            // [8 warps][2 planes][32 rows][32 bf16], dgrad only
    uint64_t* bars = reinterpret_cast<uint64_t*>(h_stage + (HTMA ? EPI_STAGE_BYTES : 0));
    uint64_t* full = bars;                          // [S] TMA -> MMA
    uint64_t* empty = full + MAX_STAGES;            // [S] MMA -> TMA
    uint64_t* tmem_full = empty + MAX_STAGES;       // [2] MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;           // [2] epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
    uint64_t* h_full = tmem_empty + 4;              // [8] TMA -> epilogue warp w: its H block has landed

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb_total = (g.K + BK - 1) / BK;
    constexpr int TILE_M = PAIR ? 2 * BM : BM;
    const int tiles_n = (g.N + g.BN - 1) / g.BN, tiles_m = (g.M + TILE_M - 1) / TILE_M;
    const int tiles_mn = tiles_n * tiles_m;
    const int total_work = tiles_mn * g.splits;
    const int S = g.stages;
    // work items are dealt round-robin to the CTAs (pair: to the clusters): neighbouring CTAs then hold the n-tiles of one m-tile at
    // the same time (the second read of those A rows is an L2 hit) and the split-K partials of one C tile are spread in time
    const int w_begin = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x, w_step = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    const int w_end = total_work;

    if (threadIdx.x == 0) {
        for (int s = 0; s < S; ++s) {
            mbar_init(&full[s], (PAIR && rank == 0) ? 2 : 1);   // pair leader: its own expect_tx arrival + the peer's remote arrival
            mbar_init(&empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], PAIR ? 2 * EPI_WARPS : EPI_WARPS);
        }
        if (DELU)
            for (int e = 0; e < EPI_WARPS; ++e) mbar_init(&h_full[e], 1);
        fence_barrier_init();
    }
    if (warp == EPI_WARPS + 1) {
        if (PAIR) tmem_alloc_2sm(tmem_slot, 512); else tmem_alloc(tmem_slot, 512);
    }
    tc_fence_before();
    if (PAIR) cluster_sync(); else __syncthreads();             // barriers initialised in BOTH CTAs before any remote arrival
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == EPI_WARPS) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            // pair, direct form: both CTAs' loads are credited to the leader's barrier (cta_group::2 TMA, a remote complete_tx per
            // arriving packet).  Relay form: every CTA's loads complete on its OWN barrier and the peer forwards ONE remote arrival.
            const bool direct = PAIR && !g.relay;
            const uint32_t tx = (uint32_t)(((g.dbg & 4) ? 0 : A_BYTES) + ((g.dbg & 8) ? 0 : 2 * b_plane)) * (direct ? 2u : 1u);
            const uint32_t full0_remote = PAIR ? mapa_u32(&full[0], 0) : 0u;
            auto load = [&](void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
                if (direct) tma_load_3d_2sm(dst, map, bar, c0, c1, 0); else tma_load_3d(dst, map, bar, c0, c1, 0);
            };
            int it = 0;
            long long t_wait = 0;
            for (int w = w_begin; w < w_end; w += w_step) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total, TILE_M);
                const int m0 = wk.m0 + (int)rank * BM, n0 = wk.n0 + (int)rank * bn_cta;       // this CTA's rows of A / rows of B
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % S, k0 = (wk.kb_begin + kb) * BK;
                    const long long tw0 = g.trace ? clock64() : 0;
                    mbar_wait(&empty[s], ((it / S) & 1) ^ 1);
                    if (g.trace) t_wait += clock64() - tw0;
                    BF3_STAMP(it, 0);
                    unsigned char* sa = smem + s * stage_bytes;
                    unsigned char* sb = sa + A_BYTES;
                    if (g.dbg & 1) {                                    // profiling: signal "loaded" without loading
                        if (!direct || rank == 0) mbar_arrive(&full[s]);
                        else mbar_arrive_remote(full0_remote + (uint32_t)(s * sizeof(uint64_t)));
                        continue;
                    }
                    if (!direct || rank == 0) mbar_expect_tx(&full[s], tx);
                    else mbar_arrive_remote(full0_remote + (uint32_t)(s * sizeof(uint64_t)));
                    if (g.dbg & 4) {}                                                             // profiling: no A loads
                    else if (!g.a_mn) load(sa, &tmA, &full[s], k0, m0);                           // [plane][128][64]
                    else
                        for (int j = 0; j < BM / 64; ++j) load(sa + j * 16384, &tmA, &full[s], m0 + 64 * j, k0);   // [plane][64 k][64 mn]
                    if (g.dbg & 8) {}                                                             // profiling: no B loads
                    else if (!g.b_mn) load(sb, &tmB, &full[s], k0, n0);                           // [plane][bn_cta][64]
                    else
                        for (int j = 0; j < bn_cta / 64; ++j) load(sb + j * 16384, &tmB, &full[s], n0 + 64 * j, k0);
                    BF3_STAMP(it, 1);
                }
            }
            if (g.trace) g.trace[(size_t)blockIdx.x * 8 + 3] = t_wait;
        }
    } else if (warp == EPI_WARPS + 1) {
        // ===== MMA issuer =====
        if (lane == 0 && rank == 0) {                           // pair: the leader issues for both CTAs
            const uint32_t idesc = make_idesc(TILE_M, g.BN, g.a_mn, g.b_mn);
            // byte offsets inside a stage: plane (hi -> lo) and K step (16 bf16)
            const uint32_t a_lo_off = g.a_mn ? 8192u : (uint32_t)A_PLANE, b_lo_off = g.b_mn ? 8192u : (uint32_t)b_plane;
            const uint32_t a_kstep = g.a_mn ? 2048u : 32u, b_kstep = g.b_mn ? 2048u : 32u;
            const uint32_t a_lbo = g.a_mn ? 16384u : 16u, b_lbo = g.b_mn ? 16384u : 16u;
            int it = 0, item = 0;
            long long t_full = 0, t_acc = 0;
            const long long t_begin = g.trace ? clock64() : 0;
            for (int w = w_begin; w < w_end; w += w_step, ++item) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total, TILE_M);
                const int acc_stage = item & 1;
                const long long ta0 = g.trace ? clock64() : 0;
                mbar_wait(&tmem_empty[acc_stage], ((item >> 1) & 1) ^ 1);       // epilogue(s) have drained this accumulator
                if (g.trace) t_acc += clock64() - ta0;
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * 256);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % S;
                    const long long tf0 = g.trace ? clock64() : 0;
                    mbar_wait(&full[s], (it / S) & 1);
                    if (g.trace) t_full += clock64() - tf0;
                    BF3_STAMP(it, 3);
                    tc_fence_after();
                    const uint32_t a0 = smem_u32(smem + s * stage_bytes), b0 = a0 + A_BYTES;
                    if (!(g.dbg & 2))
#pragma unroll
                    for (int kk = 0; kk < BK / 16; ++kk) {
                        const uint64_t a_hi = make_desc(a0 + kk * a_kstep, a_lbo);
                        const uint64_t a_lo = make_desc(a0 + a_lo_off + kk * a_kstep, a_lbo);
                        const uint64_t b_hi = make_desc(b0 + kk * b_kstep, b_lbo);
                        const uint64_t b_lo = make_desc(b0 + b_lo_off + kk * b_kstep, b_lbo);
                        if (PAIR) {
                            umma_bf16_2sm(tmem_d, a_lo, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);
                            umma_bf16_2sm(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_bf16_2sm(tmem_d, a_hi, b_hi, idesc, 1u);
                        } else {
                            umma_bf16(tmem_d, a_lo, b_hi, idesc, (kb > 0 || kk > 0) ? 1u : 0u);  // small terms first
                            umma_bf16(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_bf16(tmem_d, a_hi, b_hi, idesc, 1u);
                        }
                    }
                    if (PAIR) umma_commit_2sm(&empty[s]); else umma_commit(&empty[s]);   // stage reusable (in both CTAs) once these MMAs retire
                }
                if (PAIR) umma_commit_2sm(&tmem_full[acc_stage]); else umma_commit(&tmem_full[acc_stage]);
            }
            if (g.trace) {
                long long* tr = g.trace + (size_t)blockIdx.x * 8;
                tr[0] = clock64() - t_begin; tr[1] = t_full; tr[2] = t_acc; tr[6] = item; tr[7] = it;
            }
        } else if (PAIR && lane == 0 && rank == 1 && g.relay) {
            // ===== relay (peer CTA): "my half of stage s has landed" -> one arrival on the leader's full[s] =====
            const uint32_t full0_leader = mapa_u32(&full[0], 0);
            int it = 0;
            for (int w = w_begin; w < w_end; w += w_step) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total, TILE_M);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % S;
                    mbar_wait(&full[s], (it / S) & 1);
                    mbar_arrive_remote(full0_leader + (uint32_t)(s * sizeof(uint64_t)));
                    BF3_STAMP(it, 2);
                }
            }
        }
    } else {
        // ===== epilogue: warp w <-> TMEM lanes 32*(w%4) .. +31, columns [(w/4) * BN/2, +BN/2) =====
        // With only 8 epilogue warps every exposed latency costs: the H rows a dgrad chunk needs (ELU') are prefetched one
        // chunk ahead -- the first chunk's before the accumulator is even complete -- and bias values come as warp-uniform
        // 128-bit loads (one L1 broadcast each) instead of 32 shuffles.
        const int q = warp & 3, half = warp >> 2;
        const int c_begin = half * (g.BN >> 1), c_end = c_begin + (g.BN >> 1);
        const bool h_fast = (DELU) && ((g.ldhs & 7) == 0) && ((g.hs_plane & 7) == 0) && ((reinterpret_cast<uintptr_t>(g.Hs) & 15u) == 0);
        const bool bias_fast = (EPI == EPI_F32_BIAS || EPI == EPI_SPLIT_BIAS_ELU || EPI == EPI_DISCARD) && ((reinterpret_cast<uintptr_t>(g.bias) & 15u) == 0);
        const uint32_t tmem_empty0_leader = PAIR ? mapa_u32(&tmem_empty[0], 0) : 0u;
        int item = 0;
        long long t_wfull = 0, t_busy = 0;
        // dgrad, TMA form: the 32 x 32 x {hi, lo} block of H a chunk needs is fetched into this warp's 4 KB buffer (same SWIZZLE_64B
        // geometry as the output staging) while the previous chunk is processed -- 8 conflict-free LDS.128 per lane instead of 8
        // LDG.128 that each touch 32 different lines (256 sector requests per chunk through the L1 the tensor core saturates).
        constexpr bool h_tma = HTMA;
        unsigned char* hb = h_stage + warp * EPI_STAGE_WARP;
        uint32_t h_phase = 0;
        auto h_request = [&](int row0, int col0) {
            if (lane == 0) {
                mbar_expect_tx(&h_full[warp], (uint32_t)EPI_STAGE_WARP);
                tma_load_3d(hb, &tmH, &h_full[warp], col0, row0, 0);                   // out-of-range rows / columns arrive as zeros
            }
        };
        if (h_tma && w_begin < w_end) {
            const Work w0 = decode_work(g, w_begin, tiles_n, tiles_mn, num_kb_total, TILE_M);
            h_request(w0.m0 + (int)rank * BM + q * 32, w0.n0 + c_begin);
        }
        for (int w = w_begin; w < w_end; w += w_step, ++item) {
            const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total, TILE_M);
            const int acc_stage = item & 1;
            const long long te0 = g.trace ? clock64() : 0;
            const int row = wk.m0 + (int)rank * BM + q * 32 + lane;
            const bool row_ok = row < g.M;
            uint4 hn[8];                                                    // prefetched H (4 x hi, 4 x lo) of the NEXT chunk
            auto prefetch_h = [&](int c0) {
                const int col0 = wk.n0 + c0;
                if (h_fast && row_ok && col0 + 32 <= g.N) {
                    const uint16_t* hp = g.Hs + (int64_t)row * g.ldhs + col0;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        hn[t] = __ldg(reinterpret_cast<const uint4*>(hp) + t);
                        hn[4 + t] = __ldg(reinterpret_cast<const uint4*>(hp + g.hs_plane) + t);
                    }
                }
            };
            if (DELU && !h_tma) prefetch_h(c_begin);
            mbar_wait(&tmem_full[acc_stage], (item >> 1) & 1);
            const long long te1 = g.trace ? clock64() : 0;
            tc_fence_after();
            for (int c0 = c_begin; c0 < c_end; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc_stage * 256 + c0), v);
                const int col0 = wk.n0 + c0;
                uint4 hc[8];
                if (h_tma) {
                    mbar_wait(&h_full[warp], h_phase);
                    h_phase ^= 1u;
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int t = 0; t < 4; ++t) {
                        const unsigned char* src = hb + lane * 64 + ((t ^ sw) << 4);
                        hc[t] = *reinterpret_cast<const uint4*>(src);
                        hc[4 + t] = *reinterpret_cast<const uint4*>(src + 32 * 64);
                    }
                    __syncwarp();                                               // every lane has read the buffer: request the next block
                    if (c0 + 32 < c_end) h_request(wk.m0 + (int)rank * BM + q * 32, wk.n0 + c0 + 32);
                    else if (w + w_step < w_end) {
                        const Work nx = decode_work(g, w + w_step, tiles_n, tiles_mn, num_kb_total, TILE_M);
                        h_request(nx.m0 + (int)rank * BM + q * 32, nx.n0 + c_begin);
                    }
                } else if (DELU) {
#pragma unroll
                    for (int t = 0; t < 8; ++t) hc[t] = hn[t];
                    if (c0 + 32 < c_end) prefetch_h(c0 + 32);
                }
                if (col0 >= g.N) continue;                                  // warp-uniform
                const int nvalid = min(32, g.N - col0);
                const bool full_chunk = (nvalid == 32);
                if (EPI == EPI_F32_BIAS || EPI == EPI_SPLIT_BIAS_ELU || EPI == EPI_DISCARD) {
                    if (bias_fast && full_chunk && ((col0 & 3) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));     // same address in every lane
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
                        const float bl = (lane < nvalid) ? __ldg(g.bias + col0 + lane) : 0.0f;
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, bl, j);
                    }
                    if (EPI == EPI_SPLIT_BIAS_ELU || EPI == EPI_DISCARD) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = elu1(v[j]);
                    }
                } else if (DELU) {
                    // ELU'(z) recovered from h = ELU(z) ~= h_hi + h_lo:  1 for h > 0, h + 1 otherwise
                    if (h_tma || (h_fast && full_chunk)) {
                        if (h_tma || row_ok) {
#pragma unroll
                            for (int t = 0; t < 4; ++t) {
                                const uint32_t ah[4] = {hc[t].x, hc[t].y, hc[t].z, hc[t].w};
                                const uint32_t al[4] = {hc[4 + t].x, hc[4 + t].y, hc[4 + t].z, hc[4 + t].w};
#pragma unroll
                                for (int u = 0; u < 4; ++u) {
                                    const float h0 = bf_lo(ah[u]) + bf_lo(al[u]), h1 = bf_hi(ah[u]) + bf_hi(al[u]);
                                    v[8 * t + 2 * u] *= (h0 > 0.0f) ? 1.0f : (h0 + 1.0f);
                                    v[8 * t + 2 * u + 1] *= (h1 > 0.0f) ? 1.0f : (h1 + 1.0f);
                                }
                            }
                        }
                    } else if (row_ok) {
                        const uint16_t* hp = g.Hs + (int64_t)row * g.ldhs + col0;
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                const float h = __uint_as_float((uint32_t)hp[j] << 16) + __uint_as_float((uint32_t)hp[g.hs_plane + j] << 16);
                                v[j] *= (h > 0.0f) ? 1.0f : (h + 1.0f);
                            }
                    }
                }
                if (EPI == EPI_F32 || EPI == EPI_F32_BIAS || EPI == EPI_ATOMIC) {
                    if (!row_ok) continue;
                    float* dst = g.C + (int64_t)row * g.ldc + col0;
                    const bool vec = full_chunk && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0);
                    if (EPI == EPI_ATOMIC) {
                        if (vec) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4)
                                atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                        } else {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (j < nvalid) atomicAdd(dst + j, v[j]);
                        }
                    } else if (vec) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) dst[j] = v[j];
                    }
                    continue;
                }
                if (EPI == EPI_DISCARD) {                                 // keep the math alive, store nothing (almost)
                    float acc = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        const uint32_t ph = pack_bf16x2(v[j], v[j + 1]);
                        acc += bf_lo(pack_bf16x2(v[j] - bf_lo(ph), v[j + 1] - bf_hi(ph)));
                    }
                    if (acc == 1.2345e38f && g.Cs) g.Cs[0] = 1;
                    continue;
                }
                // ---- split store (hi / lo bf16 planes) ----
                if (g.cs_tma) {
                    // The 32 x 32 block leaves as ONE TMA store per warp (hi and lo planes): a lane's direct 16-byte stores touch 32
                    // different 128-byte lines per instruction, which crawls on an LSU/L1 path the tensor cores already saturate.
                    // Lane = row; 16-byte chunk c of a row sits at chunk c ^ ((row >> 1) & 3) (SWIZZLE_64B), so each 8-lane phase
                    // of a shared-memory store covers all 32 banks.
                    unsigned char* stg = epi_stage + warp * EPI_STAGE_WARP;
                    if (lane == 0) tma_store_wait_read();                       // the previous block's store has read the buffer
                    __syncwarp();
                    const int sw = (lane >> 1) & 3;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        uint32_t ph[4], pl[4];
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float x0 = v[8 * c + 2 * u], x1 = v[8 * c + 2 * u + 1];
                            ph[u] = pack_bf16x2(x0, x1);
                            pl[u] = pack_bf16x2(x0 - bf_lo(ph[u]), x1 - bf_hi(ph[u]));
                        }
                        unsigned char* d = stg + lane * 64 + ((c ^ sw) << 4);
                        *reinterpret_cast<uint4*>(d) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                        *reinterpret_cast<uint4*>(d + 32 * 64) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                    fence_proxy_async();                                        // generic-proxy writes -> the TMA's reads
                    __syncwarp();
                    if (lane == 0) {                                            // rows >= M and columns >= N are clipped by the map
                        tma_store_3d(&tmC, stg, col0, wk.m0 + (int)rank * BM + q * 32, 0);
                        tma_store_commit();
                    }
                } else if (row_ok) {
                    uint16_t* dp = g.Cs + (int64_t)row * g.ldcs + col0;
                    uint32_t ph[16], pl[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) {
                        ph[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
                        pl[j] = pack_bf16x2(v[2 * j] - bf_lo(ph[j]), v[2 * j + 1] - bf_hi(ph[j]));
                    }
                    if (full_chunk && ((reinterpret_cast<uintptr_t>(dp) & 15u) == 0) && ((g.cs_plane & 7) == 0)) {
#pragma unroll
                        for (int j = 0; j < 16; j += 4) {
                            *reinterpret_cast<uint4*>(dp + 2 * j) = make_uint4(ph[j], ph[j + 1], ph[j + 2], ph[j + 3]);
                            *reinterpret_cast<uint4*>(dp + g.cs_plane + 2 * j) = make_uint4(pl[j], pl[j + 1], pl[j + 2], pl[j + 3]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                dp[j] = (uint16_t)((j & 1) ? (ph[j >> 1] >> 16) : (ph[j >> 1] & 0xFFFFu));
                                dp[g.cs_plane + j] = (uint16_t)((j & 1) ? (pl[j >> 1] >> 16) : (pl[j >> 1] & 0xFFFFu));
                            }
                    }
                }
                if (DELU && g.colsum != nullptr) {
                    // bias gradient: column sums of this 32 x 32 block by a butterfly transpose-reduce (31 shuffles),
                    // lane j ends up with the sum of column j; one atomic per column per warp
                    if (!row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.0f;
                    }
#pragma unroll
                    for (int o = 16; o >= 1; o >>= 1) {
                        const bool upper = (lane & o) != 0;
#pragma unroll
                        for (int i = 0; i < o; ++i) {
                            const float send = upper ? v[i] : v[i + o];
                            const float keep = upper ? v[i + o] : v[i];
                            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
                        }
                    }
                    if (lane < nvalid) atomicAdd(g.colsum + col0 + lane, v[0]);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {                                                    // accumulator free for item + 2 (the leader's MMA thread waits)
                if (PAIR) mbar_arrive_remote(tmem_empty0_leader + (uint32_t)(acc_stage * sizeof(uint64_t)));
                else mbar_arrive(&tmem_empty[acc_stage]);
            }
            if (g.trace) { t_wfull += te1 - te0; t_busy += clock64() - te1; }
        }
        if (g.cs_tma && lane == 0) tma_store_wait_all();
        if (g.trace && threadIdx.x == 0) { g.trace[(size_t)blockIdx.x * 8 + 4] = t_wfull; g.trace[(size_t)blockIdx.x * 8 + 5] = t_busy; }
    }
    tc_fence_before();
    if (PAIR) cluster_sync(); else __syncthreads();             // the peer's shared memory stays alive until the leader's last MMA has read it
    if (warp == EPI_WARPS + 1) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_2sm(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

// ---- fp32 -> split planes (weights after every Adam step, test inputs) --------------------------------------------
__global__ void split_kernel(const float* __restrict__ src, int64_t ld_src, uint16_t* __restrict__ dst, int64_t ld_dst, int64_t plane,
                             int64_t rows, int64_t cols) {
    const int64_t pairs_per_row = (cols + 1) >> 1;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * pairs_per_row) return;
    const int64_t r = i / pairs_per_row, c = (i - r * pairs_per_row) * 2;
    const float x0 = src[r * ld_src + c], x1 = (c + 1 < cols) ? src[r * ld_src + c + 1] : 0.0f;
    const uint32_t h = pack_bf16x2(x0, x1);
    const uint32_t l = pack_bf16x2(x0 - bf_lo(h), x1 - bf_hi(h));
    uint16_t* d = dst + r * ld_dst + c;
    d[0] = (uint16_t)(h & 0xFFFFu);
    d[plane] = (uint16_t)(l & 0xFFFFu);
    if (c + 1 < ld_dst) {                                   // the pad column (c + 1 == cols < ld_dst) gets zeros
        d[1] = (uint16_t)(h >> 16);
        d[plane + 1] = (uint16_t)(l >> 16);
    }
}
__global__ void unsplit_kernel(const uint16_t* __restrict__ src, int64_t ld_src, int64_t plane, float* __restrict__ dst, int64_t ld_dst,
                               int64_t rows, int64_t cols) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= rows * cols) return;
    const int64_t r = i / cols, c = i - r * cols;
    dst[r * ld_dst + c] = __uint_as_float((uint32_t)src[r * ld_src + c] << 16) + __uint_as_float((uint32_t)src[plane + r * ld_src + c] << 16);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

int32_t load_encode() {
    if (g_encode) return 0;
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    if (e != cudaSuccess || q != cudaDriverEntryPointSuccess || !fn) return hg_fail(HG_E_STATE, "cuTensorMapEncodeTiled unavailable");
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
    return 0;
}

// Tensor maps are pure functions of (base, extents, pitches, box): the update path reuses the same scratch buffers for
// every minibatch, so each distinct map is encoded once per process and then found in this cache (~176
// host-side cuTensorMapEncodeTiled calls per update otherwise).
struct MapKey {
    const void* base; uint64_t inner, outer, ld, plane; uint32_t box_outer, box_inner;
    bool operator==(const MapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && plane == o.plane && box_outer == o.box_outer && box_inner == o.box_inner;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        const uint64_t v[6] = {(uint64_t)(uintptr_t)k.base, k.inner, k.outer, k.ld, k.plane, ((uint64_t)k.box_inner << 32) | k.box_outer};
        for (uint64_t x : v) { h ^= x; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

// 3-D bf16 map over a split tensor: {inner (contiguous), outer (rows of pitch ld), 2 planes}; box {64, box_outer, 2} with
// SWIZZLE_128B (operand loads) or {32, box_outer, 2} with SWIZZLE_64B (the epilogue's stores)
int32_t get_map(CUtensorMap* out, const uint16_t* base, uint64_t inner, uint64_t outer, uint64_t ld, uint64_t plane, uint32_t box_outer,
                uint32_t box_inner = 64) {
    MapKey key{base, inner, outer, ld, plane, box_outer, box_inner};
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *out = it->second; return 0; }
    }
    cuuint64_t dims[3] = {inner, outer, 2};
    cuuint64_t strides[2] = {ld * sizeof(uint16_t), plane * sizeof(uint16_t)};
    cuuint32_t box[3] = {box_inner, box_outer, 2};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<uint16_t*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, box_inner == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_hg_err, sizeof(g_hg_err), "cuTensorMapEncodeTiled (bf16 3-D) failed (%d): inner=%llu outer=%llu ld=%llu plane=%llu box=%u",
                 (int)r, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, (unsigned long long)plane, box_outer);
        return HG_E_ARG;
    }
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();            // bounded: callers with ever-changing pointers just re-encode
    g_maps.emplace(key, *out);
    return 0;
}

int32_t check_split(const HgSplit& s, const char* what) {
    if (!s.p) return hg_fail(HG_E_NULL, what);
    if ((s.ld & 7) || (s.plane & 7) || !hg_aligned16(s.p)) return hg_fail(HG_E_ALIGN, "hg_gemm_bf16x3: split tensors need a 16-byte aligned base, ld % 8 == 0 and plane % 8 == 0 (TMA)");
    return 0;
}

}  // namespace

template <bool PAIR>
static const void* kernel_for_pair(int epi) {
    switch (epi) {
        case EPI_F32: return (const void*)gemm_bf3_kernel<PAIR, EPI_F32>;
        case EPI_F32_BIAS: return (const void*)gemm_bf3_kernel<PAIR, EPI_F32_BIAS>;
        case EPI_SPLIT_BIAS_ELU: return (const void*)gemm_bf3_kernel<PAIR, EPI_SPLIT_BIAS_ELU>;
        case EPI_SPLIT_DELU: return (const void*)gemm_bf3_kernel<PAIR, EPI_SPLIT_DELU>;
        case EPI_ATOMIC: return (const void*)gemm_bf3_kernel<PAIR, EPI_ATOMIC>;
        case EPI_SPLIT: return (const void*)gemm_bf3_kernel<PAIR, EPI_SPLIT>;
        case EPI_DISCARD: return (const void*)gemm_bf3_kernel<PAIR, EPI_DISCARD>;
        case EPI_SPLIT_DELU_TMA: return (const void*)gemm_bf3_kernel<PAIR, EPI_SPLIT_DELU_TMA>;
    }
    return nullptr;
}
static const void* kernel_for(bool pair, int epi) { return pair ? kernel_for_pair<true>(epi) : kernel_for_pair<false>(epi); }

long long* g_bf3_trace = nullptr;
extern "C" void hg_gemm_bf16x3_set_trace(long long* buf) { g_bf3_trace = buf; }

// HG_BF3_PAIR: 0 = single-CTA form only, 1 (default) = CTA pairs (relay signalling) where they pay, 2 = same with cta_group::2 TMA
// signalling, 3 = CTA pairs (relay) wherever legal
static int bf3_pair_mode() {
    static int pair_env = -1;
    if (pair_env < 0) { const char* e = getenv("HG_BF3_PAIR"); pair_env = e ? atoi(e) : 1; }
    return pair_env;
}

extern "C" int32_t hg_split_bf16(const float* src, int64_t ld_src, const HgSplit* dst, int64_t rows, int64_t cols, void* stream) {
    HG_REQUIRE(src); HG_REQUIRE(dst); HG_REQUIRE(dst->p);
    if (rows <= 0 || cols <= 0 || ld_src < cols || dst->ld < cols) return hg_fail(HG_E_SIZE, "hg_split_bf16: bad extents");
    const int64_t n = rows * ((cols + 1) / 2);
    split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src, ld_src, dst->p, dst->ld, dst->plane, rows, cols);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_split_bf16");
}

extern "C" int32_t hg_unsplit_bf16(const HgSplit* src, float* dst, int64_t ld_dst, int64_t rows, int64_t cols, void* stream) {
    HG_REQUIRE(src); HG_REQUIRE(src->p); HG_REQUIRE(dst);
    if (rows <= 0 || cols <= 0 || ld_dst < cols || src->ld < cols) return hg_fail(HG_E_SIZE, "hg_unsplit_bf16: bad extents");
    const int64_t n = rows * cols;
    unsplit_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(src->p, src->ld, src->plane, dst, ld_dst, rows, cols);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_unsplit_bf16");
}

// C (M x N) = op(A) op(B) over K on pre-split operands, see HgGemmSplit in hg_b200.h
extern "C" int32_t hg_gemm_bf16x3(const HgGemmSplit* d, void* stream) {
    HG_REQUIRE(d);
    if (int32_t rc = check_split(d->A, "hg_gemm_bf16x3: A is NULL")) return rc;
    if (int32_t rc = check_split(d->B, "hg_gemm_bf16x3: B is NULL")) return rc;
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return hg_fail(HG_E_SIZE, "hg_gemm_bf16x3: bad extents");
    const int epi = d->epilogue;
    if (epi < 0 || epi > EPI_DISCARD) return hg_fail(HG_E_ARG, "hg_gemm_bf16x3: bad epilogue");
    const bool split_out = (epi == EPI_SPLIT_BIAS_ELU || epi == EPI_SPLIT_DELU || epi == EPI_SPLIT || epi == EPI_DISCARD);
    if (split_out) {
        if (!d->Cs.p) return hg_fail(HG_E_NULL, "hg_gemm_bf16x3: Cs is NULL");
        if (d->Cs.ld < d->N) return hg_fail(HG_E_SIZE, "hg_gemm_bf16x3: Cs.ld < N");
    } else if (!d->C) return hg_fail(HG_E_NULL, "hg_gemm_bf16x3: C is NULL");
    if ((epi == EPI_F32_BIAS || epi == EPI_SPLIT_BIAS_ELU || epi == EPI_DISCARD) && !d->bias) return hg_fail(HG_E_NULL, "hg_gemm_bf16x3: bias is NULL");
    if (epi == EPI_SPLIT_DELU && !d->Hs.p) return hg_fail(HG_E_NULL, "hg_gemm_bf16x3: Hs is NULL");
    if (int32_t rc = load_encode()) return rc;
    cudaStream_t st = (cudaStream_t)stream;

    Args g{};
    g.C = d->C; g.ldc = d->ldc;
    g.Cs = d->Cs.p; g.ldcs = d->Cs.ld; g.cs_plane = d->Cs.plane;
    g.bias = d->bias;
    g.Hs = d->Hs.p; g.ldhs = d->Hs.ld; g.hs_plane = d->Hs.plane;
    g.colsum = d->colsum;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.a_mn = d->a_mn_major ? 1 : 0; g.b_mn = d->b_mn_major ? 1 : 0;
    g.epi = epi;
    int bn = ((d->N + 63) / 64) * 64;                    // multiples of 64: MN-major boxes are 64 wide, column halves 32-aligned
    g.BN = bn > 256 ? 256 : bn;
    // CTA pairs (cta_group::2) where they measure faster (tools/bf3_trace.py): long-K launches without split-K, whose time is the
    // L2 -> SM operand feed that the pair halves for B.  Short-K launches are epilogue-bound and the pair's lock-step loads cost a
    // little, split-K launches measure equal.  The tile must be wide enough to halve (each CTA loads BN / 2 rows of B: >= 64 for
    // the MN-major boxes).  HG_BF3_PAIR=0 pins the single-CTA form, HG_BF3_PAIR=3 forces pairs wherever they are legal.
    const int pm = bf3_pair_mode();
    const bool pair_legal = g.BN >= 128 && d->M >= 256;
    const bool pair = pm != 0 && pair_legal && (pm == 3 || (d->split_k <= 1 && (d->K + BK - 1) / BK >= 8));
    // split outputs go out through shared memory and TMA when the output planes qualify for a tensor map (HG_BF3_TMA_STORE=0: direct stores)
    static int tma_store_env = -1, tma_h_env = -1;
    if (tma_store_env < 0) { const char* e = getenv("HG_BF3_TMA_STORE"); tma_store_env = (e && e[0] == '0') ? 0 : 1; }
    if (tma_h_env < 0) { const char* e = getenv("HG_BF3_TMA_H"); tma_h_env = (e && e[0] == '0') ? 0 : 1; }
    g.cs_tma = (tma_store_env && (epi == EPI_SPLIT_BIAS_ELU || epi == EPI_SPLIT_DELU || epi == EPI_SPLIT) && hg_aligned16(d->Cs.p) &&
                (d->Cs.ld & 7) == 0 && (d->Cs.plane & 7) == 0) ? 1 : 0;
    // dgrad (single-CTA form): H blocks by TMA too (HG_BF3_TMA_H=0: per-lane loads).  Their 32 KB of buffers come out of the
    // operand ring, so the tile narrows to 192 or 128 columns (whichever pads N less): these short-K launches are bound by
    // their epilogue, not by the A-tile reuse a 256-wide tile buys.
    g.h_tma = (tma_h_env && epi == EPI_SPLIT_DELU && !pair && g.cs_tma && hg_aligned16(d->Hs.p) && (d->Hs.ld & 7) == 0 && (d->Hs.plane & 7) == 0 &&
               d->Hs.ld >= d->N) ? 1 : 0;
    if (g.h_tma && d->N > 128) {
        const int w192 = (d->N + 191) / 192 * 192 - d->N, w128 = (d->N + 127) / 128 * 128 - d->N;
        g.BN = (w192 <= w128) ? 192 : 128;
    }
    const int bn_cta = pair ? g.BN / 2 : g.BN;
    const int stage_bytes = A_BYTES + 2 * bn_cta * BK * 2;
    const int epi_bytes = (g.cs_tma ? EPI_STAGE_BYTES : 0) + (g.h_tma ? EPI_STAGE_BYTES : 0);
    int stages = (SMEM_LIMIT - 1024 - BAR_BYTES - epi_bytes) / stage_bytes;
    g.stages = stages > MAX_STAGES ? MAX_STAGES : stages;
    {   // profiling knob: HG_BF3_STAGES=n caps the ring depth (tools/bf3_trace.py measures the latency it has to cover)
        static int cap = -1;
        if (cap < 0) { const char* e = getenv("HG_BF3_STAGES"); cap = e ? atoi(e) : 0; }
        if (cap > 0 && cap < g.stages) g.stages = cap;
    }
    g.trace = g_bf3_trace;
    g.relay = bf3_pair_mode() != 2;
    { static int dbg = -1; if (dbg < 0) { const char* e = getenv("HG_BF3_DEBUG"); dbg = e ? atoi(e) : 0; } g.dbg = dbg; }
    const int smem_bytes = g.stages * stage_bytes + epi_bytes + 1024 + BAR_BYTES;
    const int num_kb = (d->K + BK - 1) / BK;
    int splits = d->split_k > 0 ? d->split_k : 1;
    if (splits > num_kb) splits = num_kb;
    if (splits > 1 && epi != EPI_ATOMIC) return hg_fail(HG_E_ARG, "hg_gemm_bf16x3: split_k needs the atomic epilogue");
    g.kb_per_split = (num_kb + splits - 1) / splits;
    splits = (num_kb + g.kb_per_split - 1) / g.kb_per_split;
    g.splits = splits;

    // K-major operand: map {K, rows, 2}, box {64, tile rows, 2}.  MN-major: map {rows, K, 2}, box {64, 64, 2}.
    CUtensorMap tmA, tmB, tmC, tmH;
    memset(&tmC, 0, sizeof(tmC));
    memset(&tmH, 0, sizeof(tmH));
    int32_t rc;
    if (g.h_tma) {
        rc = get_map(&tmH, d->Hs.p, d->N, d->M, d->Hs.ld, d->Hs.plane, 32, 32);
        if (rc) return rc;
    }
    if (g.cs_tma) {
        rc = get_map(&tmC, d->Cs.p, d->N, d->M, d->Cs.ld, d->Cs.plane, 32, 32);
        if (rc) return rc;
    }
    if (!g.a_mn) rc = get_map(&tmA, d->A.p, d->K, d->M, d->A.ld, d->A.plane, BM);
    else rc = get_map(&tmA, d->A.p, d->M, d->K, d->A.ld, d->A.plane, BK);
    if (rc) return rc;
    if (!g.b_mn) rc = get_map(&tmB, d->B.p, d->K, d->N, d->B.ld, d->B.plane, bn_cta);
    else rc = get_map(&tmB, d->B.p, d->N, d->K, d->B.ld, d->B.plane, BK);
    if (rc) return rc;

    const int tile_m = pair ? 2 * BM : BM;
    const int total_work = ((d->N + g.BN - 1) / g.BN) * ((d->M + tile_m - 1) / tile_m) * splits;
    const int workers = pair ? HG_NUM_SMS / 2 : HG_NUM_SMS;
    const int grid = (total_work < workers ? total_work : workers) * (pair ? 2 : 1);          // persistent: one CTA (pair: one cluster) per SM (pair)
    const int kepi = g.h_tma ? (int)EPI_SPLIT_DELU_TMA : epi;
    const void* fn = kernel_for(pair, kepi);
    if (!fn) return hg_fail(HG_E_ARG, "hg_gemm_bf16x3: bad epilogue");
    {
        static bool attr_set[64][2][EPI_SPLIT_DELU_TMA + 1] = {};
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev >= 0 && dev < 64 && !attr_set[dev][pair ? 1 : 0][kepi]) {
            cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT);
            if (e != cudaSuccess) return hg_fail((int32_t)e, cudaGetErrorString(e));
            attr_set[dev][pair ? 1 : 0][kepi] = true;
        }
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem_bytes; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = pair ? 1 : 0;
    void* kargs[5] = {&tmA, &tmB, &tmC, &tmH, &g};
    cudaError_t e = cudaLaunchKernelExC(&cfg, fn, kargs);
    if (e != cudaSuccess) return hg_fail((int32_t)e, cudaGetErrorString(e));
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_gemm_bf16x3");
}
