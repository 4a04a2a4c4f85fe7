// tcgen05 / TMEM / TMA GEMM for the ActorCritic MLP (sm_100a), fp32 in / fp32 out with 3xTF32
// split-compensation:   D = A_hi B_hi + A_lo B_hi + A_hi B_lo,   x_hi = tf32(x) (top 19 bits),  x_lo = x - x_hi
// which restores ~fp32 accuracy (gradient rel. error ~7e-7) while every multiply runs
// on the 5th-generation tensor cores.  `passes = 1` gives plain TF32.
//
// One kernel covers the three products of a Linear layer, without materialising any transpose:
//     forward  Y  = X  W^T   A = X  (K-major)    B = W  (K-major)
//     dgrad    dX = dZ W     A = dZ (K-major)    B = W  (MN-major: the GEMM's N index is W's contiguous dim)
//     wgrad    dW = dZ^T X   A = dZ (MN-major)   B = X  (MN-major), split-K over the batch, fp32 atomics
// (UMMA instruction-descriptor bits 15/16 select K- vs MN-major per operand.)
//
// CTA = 128 x BN output tile, 14 warps, warp-specialised:
//     warps 0-3  epilogue     tcgen05.ld 32 TMEM lanes each -> bias / ELU / ELU' / atomics -> global
//     warp  4    TMA producer cp.async.bulk.tensor (128B-swizzled 128x32 fp32 tiles, OOB zero fill)
//     warp  5    MMA issuer   one lane issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8), accumulator in TMEM
//     warps 6-13 splitter     raw fp32 tile -> lo tile (x - tf32(x)), elementwise so the swizzle is untouched
// 4-deep ring of raw {A, B} tile pairs (TMA look-ahead) + 2-deep ring of {A_lo, B_lo} pairs, mbarrier-linked; CTAs are
// persistent (one per SM) and the accumulator is double-buffered in TMEM so that epilogue and main loop overlap.
#include <cuda.h>

#include <mutex>
#include <unordered_map>

#include "hg_common.cuh"
#include "hg_tc_ptx.cuh"

namespace {

constexpr int BM = 128, BK = 32;
constexpr int RAW_STAGES = 4;                        // TMA look-ahead: raw {A, B} tiles
constexpr int LO_STAGES = 2;                         // splitter -> MMA: {A_lo, B_lo} tiles
constexpr int TILE_BYTES = BM * BK * 4;              // 16 KB: 128 rows (or 4 MN-boxes) x 128 B
constexpr int PAIR_BYTES = 2 * TILE_BYTES;           // one {A, B} pair
constexpr int SPLIT_WARPS = 8;
constexpr int TC_THREADS = (6 + SPLIT_WARPS) * 32;   // 4 epilogue + TMA + MMA + splitters
constexpr int SMEM_BYTES = (RAW_STAGES + LO_STAGES) * PAIR_BYTES + 1024 /*align*/ + 256 /*barriers*/;      // upper bound over all configurations

enum { EPI_STORE = 0, EPI_BIAS = 1, EPI_BIAS_ELU = 2, EPI_MUL_DELU = 3, EPI_ATOMIC = 4, EPI_BIAS_SAMPLE = 5 };
constexpr float kLogSqrt2Pi = 0.9189385332046727f;

struct TcArgs {
    float* C; const float* bias; const float* H;
    int M, N, K;                     // GEMM extents: C is M x N, reduction over K
    int64_t ldc, ldh;
    int BN;                          // 32 / 64 / 96 / 128
    int a_mn, b_mn;                  // operand majorness (0 = K-major, 1 = MN-major)
    int epi, passes;
    int kb_per_split;                // k-blocks (of 32) per split
    int splits;
    int hi_in_place;                 // 1: splitter rewrites the raw tile with its tf32 truncation
    int b_lo_tma;                    // 1: B_lo (pre-split weights) arrives by TMA inside the raw stage; the splitter handles A only
    // EPI_BIAS_SAMPLE (PPO.act fused into the output layer, ppo.py:91-101): C receives the mean
    const float* stdv; const float* eps; float* actions; float* logp; float* sigma;
    uint64_t seed, step; const uint64_t* step_dev;
};

using namespace hgtc;

// UMMA shared-memory matrix descriptor, descriptor version 1 (sm_100).
//   K-major  tile [rows][32 fp32], SWIZZLE_128B (type 2)          : SBO = 1024 B (8 rows x 128 B), LBO unused (1)
//   MN-major tile 4 boxes of [32 k][32 fp32], 32-bit operands must use SWIZZLE_128B_BASE32B (type 1: 32-byte
//            chunks swizzled within 128 B, pattern period 4 rows)  : SBO = 512 B (4 k-rows), LBO = 4096 B (next 32 along MN)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, bool mn_major) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(mn_major ? (4096 >> 4) : 1) << 16;
    d |= (uint64_t)((mn_major ? 512 : 1024) >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(mn_major ? 1 : 2) << 61;
    return d;
}
// instruction descriptor: D = F32, A = B = TF32, M = 128
__device__ __forceinline__ uint32_t make_idesc(int N, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                       // c_format F32
    d |= 2u << 7;                       // a_format TF32
    d |= 2u << 10;                      // b_format TF32
    d |= (a_mn ? 1u : 0u) << 15;
    d |= (b_mn ? 1u : 0u) << 16;
    d |= (uint32_t)(N >> 3) << 17;
    d |= (uint32_t)(BM >> 4) << 24;
    return d;
}

// Work item w = (m_tile, n_tile, split).  CTAs are persistent: CTA c processes items c, c + gridDim.x, ...;
// the {A,B,A_lo,B_lo} stage ring keeps rolling across items and the accumulator is double-buffered in TMEM
// (2 x 128 columns), so the epilogue of item i overlaps the main loop of item i + 1.
struct Work { int m0, n0, kb_begin, num_kb; };
__device__ __forceinline__ Work decode_work(const TcArgs& g, int w, int tiles_n, int tiles_mn, int num_kb_total) {
    Work r;
    const int split = w / tiles_mn, t = w - split * tiles_mn;
    const int tm = t / tiles_n, tn = t - tm * tiles_n;
    r.m0 = tm * BM;
    r.n0 = tn * g.BN;
    r.kb_begin = split * g.kb_per_split;
    r.num_kb = min(num_kb_total, r.kb_begin + g.kb_per_split) - r.kb_begin;
    return r;
}

__global__ void __launch_bounds__(TC_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
               const TcArgs g) {
    extern __shared__ unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    // raw stage = {A, B [, B_lo]}, lo stage = {A_lo [, B_lo]}: with pre-split weights (b_lo_tma) B_lo rides in the raw stage
    const int b_bytes = g.BN * BK * 4;
    const int raw_stride = TILE_BYTES + b_bytes + (g.b_lo_tma ? b_bytes : 0);
    const int lo_stride = TILE_BYTES + (g.b_lo_tma ? 0 : b_bytes);
    unsigned char* smem_lo = smem + RAW_STAGES * raw_stride;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_lo + LO_STAGES * lo_stride);
    uint64_t* full = bars;                          // [RAW] TMA -> splitter (and MMA)
    uint64_t* empty = full + RAW_STAGES;            // [RAW] MMA -> TMA
    uint64_t* ready = empty + RAW_STAGES;           // [LO]  splitter -> MMA
    uint64_t* lo_empty = ready + LO_STAGES;         // [LO]  MMA -> splitter
    uint64_t* tmem_full = lo_empty + LO_STAGES;     // [2]   MMA -> epilogue
    uint64_t* tmem_empty = tmem_full + 2;           // [2]   epilogue -> MMA
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int num_kb_total = (g.K + BK - 1) / BK;
    const int tiles_n = (g.N + g.BN - 1) / g.BN, tiles_m = (g.M + BM - 1) / BM;
    const int tiles_mn = tiles_n * tiles_m;
    const int total_work = tiles_mn * g.splits;

    if (threadIdx.x == 0) {
        for (int s = 0; s < RAW_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
        }
        for (int s = 0; s < LO_STAGES; ++s) {
            mbar_init(&ready[s], SPLIT_WARPS);
            mbar_init(&lo_empty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&tmem_full[a], 1);
            mbar_init(&tmem_empty[a], 4);
        }
        fence_barrier_init();
    }
    if (warp == 5) tmem_alloc(tmem_slot, 256);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            if (g.b_lo_tma) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
            const uint32_t tx = TILE_BYTES + (uint32_t)b_bytes * (g.b_lo_tma ? 2u : 1u);
            int it = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % RAW_STAGES, k0 = (wk.kb_begin + kb) * BK;
                    mbar_wait(&empty[s], ((it / RAW_STAGES) & 1) ^ 1);
                    unsigned char* st = smem + s * raw_stride;
                    mbar_expect_tx(&full[s], tx);
                    if (g.b_lo_tma) tma_load_2d(st + TILE_BYTES + b_bytes, &tmBlo, &full[s], k0, wk.n0);      // K-major weights only
                    if (!g.a_mn) tma_load_2d(st, &tmA, &full[s], k0, wk.m0);
                    else
                        for (int j = 0; j < BM / 32; ++j) tma_load_2d(st + j * 4096, &tmA, &full[s], wk.m0 + 32 * j, k0);
                    if (!g.b_mn) tma_load_2d(st + TILE_BYTES, &tmB, &full[s], k0, wk.n0);
                    else
                        for (int j = 0; j < g.BN / 32; ++j) tma_load_2d(st + TILE_BYTES + j * 4096, &tmB, &full[s], wk.n0 + 32 * j, k0);
                }
            }
        }
    } else if (warp == 5) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(g.BN, g.a_mn, g.b_mn);
            const uint32_t kstep_a = g.a_mn ? 1024 : 32, kstep_b = g.b_mn ? 1024 : 32;   // bytes per K = 8
            int it = 0, item = 0;
            for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
                const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
                const int acc_stage = item & 1;
                mbar_wait(&tmem_empty[acc_stage], ((item >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + (uint32_t)(acc_stage * 128);
                for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                    const int s = it % RAW_STAGES, l = it % LO_STAGES;
                    mbar_wait(&ready[l], (it / LO_STAGES) & 1);                   // lo tiles written (raw tiles landed before that)
                    mbar_wait(&full[s], (it / RAW_STAGES) & 1);
                    tc_fence_after();
                    const uint32_t base = smem_u32(smem + s * raw_stride), base_lo = smem_u32(smem_lo + l * lo_stride);
                    const uint32_t b_lo_base = g.b_lo_tma ? base + TILE_BYTES + b_bytes : base_lo + TILE_BYTES;
#pragma unroll
                    for (int kk = 0; kk < BK / 8; ++kk) {
                        const uint64_t a_hi = make_desc(base + kk * kstep_a, g.a_mn);
                        const uint64_t b_hi = make_desc(base + TILE_BYTES + kk * kstep_b, g.b_mn);
                        const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u;
                        if (g.passes == 3) {
                            const uint64_t a_lo = make_desc(base_lo + kk * kstep_a, g.a_mn);
                            const uint64_t b_lo = make_desc(b_lo_base + kk * kstep_b, g.b_mn);
                            umma_tf32(tmem_d, a_lo, b_hi, idesc, acc);          // small terms first
                            umma_tf32(tmem_d, a_hi, b_lo, idesc, 1u);
                            umma_tf32(tmem_d, a_hi, b_hi, idesc, 1u);
                        } else {
                            umma_tf32(tmem_d, a_hi, b_hi, idesc, acc);
                        }
                    }
                    umma_commit(&empty[s]);                                     // stages reusable once these MMAs retire
                    umma_commit(&lo_empty[l]);
                }
                umma_commit(&tmem_full[acc_stage]);
            }
        }
    } else if (warp >= 6) {
        // ===== splitter: x -> (tf32(x), x - tf32(x)), elementwise (swizzle-agnostic) =====
        const int t = threadIdx.x - 6 * 32;                                     // 0 .. 32*SPLIT_WARPS-1
        constexpr int NT = 32 * SPLIT_WARPS;
        const int nB4 = g.BN * BK / 4;                                          // float4 count of the B tile
        int it = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
            const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
            for (int kb = 0; kb < wk.num_kb; ++kb, ++it) {
                const int s = it % RAW_STAGES, l = it % LO_STAGES;
                mbar_wait(&lo_empty[l], ((it / LO_STAGES) & 1) ^ 1);            // MMA is done with this lo pair
                mbar_wait(&full[s], (it / RAW_STAGES) & 1);
                if (g.passes == 3) {
                    float4* a = reinterpret_cast<float4*>(smem + s * raw_stride);
                    float4* b = a + TILE_BYTES / 16;
                    float4* alo = reinterpret_cast<float4*>(smem_lo + l * lo_stride);
                    float4* blo = alo + TILE_BYTES / 16;
                    auto split = [&](float4* raw, float4* lo, int i) {
                        // hi = truncation (what the tensor core does to a raw fp32 operand anyway);
                        // lo = residual, rounded to nearest tf32 so that its own truncation error vanishes
                        float4 x = raw[i], h, l;
                        h.x = __uint_as_float(__float_as_uint(x.x) & 0xFFFFE000u); l.x = rna_tf32(x.x - h.x);
                        h.y = __uint_as_float(__float_as_uint(x.y) & 0xFFFFE000u); l.y = rna_tf32(x.y - h.y);
                        h.z = __uint_as_float(__float_as_uint(x.z) & 0xFFFFE000u); l.z = rna_tf32(x.z - h.z);
                        h.w = __uint_as_float(__float_as_uint(x.w) & 0xFFFFE000u); l.w = rna_tf32(x.w - h.w);
                        lo[i] = l;
                        if (g.hi_in_place) raw[i] = h;
                    };
#pragma unroll 4
                    for (int i = t; i < TILE_BYTES / 16; i += NT) split(a, alo, i);
                    if (!g.b_lo_tma) {
#pragma unroll 4
                        for (int i = t; i < nB4; i += NT) split(b, blo, i);
                    }
                    fence_proxy_async();                                        // generic-proxy writes -> tensor-core reads
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[l]);
            }
        }
    } else {
        // ===== epilogue (warps 0-3 <-> TMEM lanes 32*warp .. +31) =====
        int item = 0;
        for (int w = blockIdx.x; w < total_work; w += gridDim.x, ++item) {
            const Work wk = decode_work(g, w, tiles_n, tiles_mn, num_kb_total);
            const int acc_stage = item & 1;
            mbar_wait(&tmem_full[acc_stage], (item >> 1) & 1);
            tc_fence_after();
            const int row = wk.m0 + warp * 32 + lane;
            const bool row_ok = row < g.M;
            for (int c0 = 0; c0 < g.BN; c0 += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)(acc_stage * 128 + c0), v);
                const int col0 = wk.n0 + c0;
                if (col0 >= g.N) continue;                                  // warp-uniform
                float* dst = g.C + (int64_t)row * g.ldc + col0;
                const int nvalid = min(32, g.N - col0);
                const bool vec_ok = (nvalid == 32);
                if (g.epi == EPI_BIAS || g.epi == EPI_BIAS_ELU || g.epi == EPI_BIAS_SAMPLE) {
                    // bias: warp-uniform loads (L1 broadcast), not shuffles -- the MIO pipe is saturated by tensor-core operand reads
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(g.bias + col0) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 b4 = __ldg(reinterpret_cast<const float4*>(g.bias + col0 + j));
                            v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) v[j] += __ldg(g.bias + col0 + j);
                    }
                    if (g.epi == EPI_BIAS_ELU) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = elu_fp32(v[j]);                                  // nn.ELU(alpha=1)
                    }
                } else if (g.epi == EPI_MUL_DELU) {
                    const float* h = g.H + (int64_t)row * g.ldh + col0;
                    if (row_ok && vec_ok && ((reinterpret_cast<uintptr_t>(h) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4) {
                            const float4 hv = __ldg(reinterpret_cast<const float4*>(h + j));
                            v[j] *= (hv.x > 0.0f) ? 1.0f : (hv.x + 1.0f);
                            v[j + 1] *= (hv.y > 0.0f) ? 1.0f : (hv.y + 1.0f);
                            v[j + 2] *= (hv.z > 0.0f) ? 1.0f : (hv.z + 1.0f);
                            v[j + 3] *= (hv.w > 0.0f) ? 1.0f : (hv.w + 1.0f);
                        }
                    } else if (row_ok) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) {
                                float hv = __ldg(h + j);
                                v[j] *= (hv > 0.0f) ? 1.0f : (hv + 1.0f);
                            }
                    }
                }
                if (!row_ok) continue;
                if (g.epi == EPI_BIAS_SAMPLE) {
                    // ActorCritic.act + get_actions_log_prob (actor_critic.py:111-120) on the row this thread owns:
                    // a = mu + sigma z, log-prob summed over the actions, sigma broadcast (same arithmetic as policy_sample_kernel)
                    uint64_t stp = g.step_dev ? *g.step_dev : g.step;
                    float lp = 0.0f;
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nvalid) {
                            const float mu = v[j];
                            const float sg = mu * 0.0f + __ldg(g.stdv + j);
                            float z;
                            if (g.eps) z = g.eps[(size_t)row * g.N + j];
                            else {
                                HgPhilox r = hg_philox(g.seed, (uint32_t)row, (uint32_t)stp, HG_RNG_SAMPLE | ((uint32_t)(stp >> 32) << 8), j);
                                z = hg_normal(r.c[0], r.c[1]);
                            }
                            const float a = mu + sg * z;
                            const float d = a - mu;
                            lp += -(d * d) / (2.0f * (sg * sg)) - logf(sg) - kLogSqrt2Pi;
                            dst[j] = mu;
                            g.actions[(size_t)row * g.N + j] = a;
                            g.sigma[(size_t)row * g.N + j] = sg;
                        }
                    g.logp[row] = lp;
                    continue;
                }
                if (g.epi == EPI_ATOMIC) {
                    if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                        for (int j = 0; j < 32; j += 4)
                            atomicAdd(reinterpret_cast<float4*>(dst + j), make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]));
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (j < nvalid) atomicAdd(dst + j, v[j]);
                    }
                } else if (vec_ok && ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (j < nvalid) dst[j] = v[j];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[acc_stage]);                 // accumulator free for item + 2
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 256);
    }
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

// 2-D fp32 tensor map: `inner` contiguous elements, `outer` rows of pitch `ld` elements; 128B-swizzled boxes
int32_t make_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer,
                 bool mn_major) {
    cuuint64_t dims[2] = {inner, outer};
    cuuint64_t strides[1] = {ld * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_outer};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = g_encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, mn_major ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        snprintf(g_hg_err, sizeof(g_hg_err), "cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                 (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld);
        return HG_E_ARG;
    }
    return 0;
}

// Tensor maps are pure functions of (base, extents, pitch, box): encode each distinct one once per process.
struct MapKey {
    const void* base; uint64_t inner, outer, ld; uint32_t box_inner, box_outer; int mn;
    bool operator==(const MapKey& o) const {
        return base == o.base && inner == o.inner && outer == o.outer && ld == o.ld && box_inner == o.box_inner && box_outer == o.box_outer && mn == o.mn;
    }
};
struct MapKeyHash {
    size_t operator()(const MapKey& k) const {
        uint64_t h = 1469598103934665603ull;
        const uint64_t v[7] = {(uint64_t)(uintptr_t)k.base, k.inner, k.outer, k.ld, k.box_inner, k.box_outer, (uint64_t)k.mn};
        for (uint64_t x : v) { h ^= x; h *= 1099511628211ull; }
        return (size_t)h;
    }
};
std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
std::mutex g_maps_mu;

int32_t get_map(CUtensorMap* map, const float* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer, bool mn) {
    MapKey key{base, inner, outer, ld, box_inner, box_outer, mn ? 1 : 0};
    {
        std::lock_guard<std::mutex> lk(g_maps_mu);
        auto it = g_maps.find(key);
        if (it != g_maps.end()) { *map = it->second; return 0; }
    }
    if (int32_t rc = make_map(map, base, inner, outer, ld, box_inner, box_outer, mn)) return rc;
    std::lock_guard<std::mutex> lk(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *map);
    return 0;
}

}  // namespace

// C (M x N) = op(A) op(B) over K, see HgGemm in hg_b200.h
extern "C" int32_t hg_gemm_tf32(const HgGemm* d, void* stream) {
    HG_REQUIRE(d); HG_REQUIRE(d->A); HG_REQUIRE(d->B); HG_REQUIRE(d->C);
    if (d->M <= 0 || d->N <= 0 || d->K <= 0) return hg_fail(HG_E_SIZE, "hg_gemm_tf32: bad extents");
    if (d->passes != 1 && d->passes != 3) return hg_fail(HG_E_ARG, "hg_gemm_tf32: passes must be 1 (plain TF32) or 3 (3xTF32)");
    if ((d->lda & 3) || (d->ldb & 3) || !hg_aligned16(d->A) || !hg_aligned16(d->B))
        return hg_fail(HG_E_ALIGN, "hg_gemm_tf32: operands need 16-byte aligned base and row pitch (TMA)");
    const int epi = d->epilogue;
    if (epi < 0 || epi > EPI_BIAS_SAMPLE) return hg_fail(HG_E_ARG, "hg_gemm_tf32: bad epilogue");
    if ((epi == EPI_BIAS || epi == EPI_BIAS_ELU || epi == EPI_BIAS_SAMPLE) && !d->bias) return hg_fail(HG_E_NULL, "hg_gemm_tf32: bias is NULL");
    if (epi == EPI_MUL_DELU && !d->H) return hg_fail(HG_E_NULL, "hg_gemm_tf32: H is NULL");
    if (epi == EPI_BIAS_SAMPLE) {
        if (d->N > 32) return hg_fail(HG_E_SIZE, "hg_gemm_tf32: the fused sampling epilogue needs N <= 32");
        if (!d->sample_std || !d->sample_actions || !d->sample_log_prob || !d->sample_sigma) return hg_fail(HG_E_NULL, "hg_gemm_tf32: sampling outputs are NULL");
    }
    const bool b_lo = d->B_lo != nullptr && d->passes == 3 && !d->b_mn_major;
    if (d->B_lo && !hg_aligned16(d->B_lo)) return hg_fail(HG_E_ALIGN, "hg_gemm_tf32: B_lo must be 16-byte aligned");
    if (int32_t rc = load_encode()) return rc;
    cudaStream_t st = (cudaStream_t)stream;

    TcArgs g{};
    g.C = d->C; g.bias = d->bias; g.H = d->H;
    g.M = d->M; g.N = d->N; g.K = d->K;
    g.ldc = d->ldc; g.ldh = d->ldh;
    g.a_mn = d->a_mn_major ? 1 : 0; g.b_mn = d->b_mn_major ? 1 : 0;
    g.epi = epi; g.passes = d->passes;
    g.hi_in_place = d->trust_hw_truncation ? 0 : 1;
    g.b_lo_tma = b_lo ? 1 : 0;
    g.stdv = d->sample_std; g.eps = d->sample_eps; g.actions = d->sample_actions; g.logp = d->sample_log_prob; g.sigma = d->sample_sigma;
    g.seed = d->sample_seed; g.step = d->sample_step; g.step_dev = d->sample_step_dev;
    // Tile width: 128 columns when there are enough row tiles to fill the chip, narrower for skinny batches (the
    // rollout's M = 4096 gives only 32 row tiles: BN = 64 / 32 spreads a layer over 4x more SMs and the per-CTA
    // main loop -- which is what such a launch waits for -- shrinks with it)
    const int tiles_m = (d->M + BM - 1) / BM;
    int bn = ((d->N + 31) / 32) * 32;
    if (bn > 128) bn = 128;
    if (epi != EPI_ATOMIC && d->split_k <= 1)
        while (bn > 32 && tiles_m * ((d->N + bn - 1) / bn) < HG_NUM_SMS * 3 / 4 && (bn / 2) % 32 == 0) bn /= 2;
    if (epi == EPI_BIAS_SAMPLE) bn = 32;
    g.BN = bn;
    const int num_kb = (d->K + BK - 1) / BK;
    int splits = d->split_k > 0 ? d->split_k : 1;
    if (splits > num_kb) splits = num_kb;
    if (splits > 1 && d->epilogue != EPI_ATOMIC) return hg_fail(HG_E_ARG, "hg_gemm_tf32: split_k needs the atomic epilogue");
    g.kb_per_split = (num_kb + splits - 1) / splits;
    splits = (num_kb + g.kb_per_split - 1) / g.kb_per_split;

    // K-major operand: tensor (inner = K, outer = rows), box {32, rows}.  MN-major: (inner = rows, outer = K), box {32, 32}.
    CUtensorMap tmA, tmB, tmBlo;
    int32_t rc;
    if (!g.a_mn) rc = get_map(&tmA, d->A, d->K, d->M, d->lda, BK, BM, false);
    else rc = get_map(&tmA, d->A, d->M, d->K, d->lda, 32, BK, true);
    if (rc) return rc;
    if (!g.b_mn) rc = get_map(&tmB, d->B, d->K, d->N, d->ldb, BK, g.BN, false);
    else rc = get_map(&tmB, d->B, d->N, d->K, d->ldb, 32, BK, true);
    if (rc) return rc;
    if (b_lo) {
        if (int32_t rc2 = get_map(&tmBlo, d->B_lo, d->K, d->N, d->ldb, BK, g.BN, false)) return rc2;
    } else tmBlo = tmB;

    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        if (e != cudaSuccess) return hg_fail((int32_t)e, cudaGetErrorString(e));
        attr_set[dev] = true;
    }
    g.splits = splits;
    const int total_work = ((d->N + g.BN - 1) / g.BN) * tiles_m * splits;
    const int grid = total_work < HG_NUM_SMS ? total_work : HG_NUM_SMS;       // persistent: one CTA per SM
    const int b_bytes = g.BN * BK * 4;
    const int smem_bytes = RAW_STAGES * (TILE_BYTES + b_bytes * (b_lo ? 2 : 1)) + LO_STAGES * (TILE_BYTES + (b_lo ? 0 : b_bytes)) + 1024 + 256;
    gemm_tc_kernel<<<grid, TC_THREADS, smem_bytes, st>>>(tmA, tmB, tmBlo, g);
    HG_LAUNCHED(1);
    return hg_cuda_status("hg_gemm_tf32");
}
